"""The nine legacy entry points (NNSPClass_*, FeatureClass_*, NeuralNetClass_*) on the GPU.

oracle/_ref/libnnsp_dropin.so is the reference's OWN controller (evb/src/nnCntrlClass.c, PcmBufClass.c) and
model tables (evb/src/def_nn*.c), compiled unmodified against include/nnsp_compat and linked to
libnnsp_b200.so in place of ns-nnsp.a. Driving it must give exactly what the CPU oracle gives."""
import numpy as np
import pytest

from oracle.pyoracle import RefLib

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not RefLib.available(dropin=True), reason="oracle/_ref/libnnsp_dropin.so not built")]


@pytest.fixture(scope="module")
def dropin():
    return RefLib(dropin=True)


@pytest.mark.parametrize("nn_id", [0, 1, 2])
def test_legacy_nnspclass_exec_every_tap(nb, oracle, dropin, nn_id):
    """NNSPClass_init/_reset/_exec + the debug_layer tap of NeuralNetClass_exe, frame by frame."""
    pcm = nb.synth_pcm(6, 120, first_stream=4)
    m = oracle.model(nn_id, False)
    n0 = nb.kernel_launches()
    for row in pcm[[0, 1, 4, 5]]:
        r_gpu, t_gpu = dropin.nnsp_run(nn_id, row)
        r_or, t_or = oracle.nnsp_run(m, row)
        assert (r_gpu == r_or).all()
        for name in t_or.names():
            assert (getattr(t_gpu, name) == getattr(t_or, name)).all(), name
    assert nb.kernel_launches() > n0          # the legacy symbols really ran CUDA kernels


def test_legacy_reset_semantics(nb, oracle, dropin):
    """NNSPClass_reset on a live instance keeps context row 5 (feature_module.c:39-42)."""
    pcm = nb.synth_pcm(1, 80, first_stream=9)[0]
    m = oracle.model(1, False)
    st = oracle.lib.nnsp_oracle_stream_new()
    oracle.nnsp_run(m, pcm[:40 * 160], state=st, reset=1, taps=False)
    r_or, t_or = oracle.nnsp_run(m, pcm[40 * 160:], state=st, reset=2)
    oracle.lib.nnsp_oracle_stream_free(st)
    dropin.nnsp_run(1, pcm[:40 * 160], reset=1, taps=False)
    r_gpu, t_gpu = dropin.nnsp_run(1, pcm[40 * 160:], reset=2)
    assert (r_gpu == r_or).all() and (t_gpu.act == t_or.act).all() and (t_gpu.c == t_or.c).all()


def test_reference_controller_on_the_cuda_engine(nb, oracle, dropin):
    """nnCntrlClass_exec (reference source) -> NNSPClass_exec (CUDA): stage walk identical to the oracle."""
    pcm = nb.synth_pcm(6, 700, first_stream=4)
    om = [oracle.model(i, False) for i in range(3)]
    for row in pcm[[1, 5]]:
        r_gpu, t_gpu, v_gpu = dropin.cascade_run(row)
        r_or, t_or, v_or = oracle.cascade_run(om, row)
        for f in r_or.dtype.names:
            if not (r_gpu[f] == r_or[f]).all():
                t = int(np.nonzero((r_gpu[f] != r_or[f]).reshape(len(r_or), -1).any(axis=1))[0][0])
                raise AssertionError("%s differs first at frame %d: gpu %s oracle %s" % (f, t, r_gpu[t], r_or[t]))
        assert (v_gpu == v_or).all() and (t_gpu.feat == t_or.feat).all() and (t_gpu.c == t_or.c).all()
