"""Adversarial network inputs ON the CUDA kernels: NeuralNetClass_exe (neural_nets.c:44-168) evaluated on explicit
(input, LSTM state) pairs by every network path (dp2a warp-per-stream, IMMA in the time loop, scan-split) through
nnsp_b200_net_eval, against fixtures generated from the unmodified reference:
  * golden_v1.npz `net_*`: the three shipped models, both accumulator modes -- full-scale x, extreme h / c;
    this attacks the load-time proof (MmaLayer.fast) that no clamp of affine.c:190-249 can fire on the default path;
  * golden_nets_v1.npz: synthetic and crafted stacks -- shift_32b saturation (affine_acc32b.c:566-592), bias << 30 wrap
    (affine_acc32b.c:190-217), tanh_fix past 5.0 (activation.c:31-69), cell state sat32 (lstm.c:106-115).
The same shipped-model vectors also go through the LEGACY NeuralNetClass_exe symbol (reference glue linked to the CUDA
library, oracle/_ref/libnnsp_dropin.so)."""
import numpy as np
import pytest

from common import GOLDEN_NETS, MODEL_NAME, NET_CASES, NET_CRAFTED, golden, net_case_blob, net_case_dims
from oracle.pyoracle import RefLib

pytestmark = pytest.mark.gpu
G = golden()
MODEL_FILE = {0: "s2i.nnspm", 1: "vad.nnspm", 2: "kws_galaxy.nnspm"}
PATHS = ["split", "imma", "dp2a"]


@pytest.mark.parametrize("acc32", [False, True])
@pytest.mark.parametrize("nn_id", [0, 1, 2])
def test_shipped_models_adversarial_vectors_every_path(nb, nn_id, acc32):
    m = nb.Model.from_blob(nb.MODEL_DIR + "/" + MODEL_FILE[nn_id], acc32=acc32)
    key = "net_%s_%s" % (MODEL_NAME[nn_id], "acc32" if acc32 else "acc64")
    n0 = nb.kernel_launches()
    for path in PATHS:
        act, logits, h1, c1 = nb.net_eval(m, G["net_x"], G[key + "_h0"], G[key + "_c0"], nn_path=path)
        for got, name in ((act, "_act"), (logits, "_logits"), (h1, "_h1"), (c1, "_c1")):
            bad = np.nonzero((got != G[key + name]).reshape(len(got), -1).any(axis=1))[0]
            assert len(bad) == 0, "%s path %s: %s differs on vectors %s" % (key, path, name, bad.tolist())
    assert nb.kernel_launches() > n0


@pytest.mark.parametrize("acc32", [False, True])
@pytest.mark.parametrize("case", NET_CASES + NET_CRAFTED, ids=[c[0] for c in NET_CASES + NET_CRAFTED])
def test_synthetic_and_crafted_networks_every_path(nb, case, acc32):
    N = np.load(GOLDEN_NETS)
    blob = net_case_blob(case, acc32)
    key = "%s_%s" % (case[0], "acc32" if acc32 else "acc64")
    a_s, h_s, n_o = net_case_dims(case)
    m = nb.Model.from_blob(blob, acc32=acc32)
    ran = 0
    for path in ["auto"] + PATHS:
        try:
            act, logits, h1, c1 = nb.net_eval(m, N["x"], N[key + "_h0"], N[key + "_c0"], nn_path=path)
        except nb.NnspError as e:
            assert path in ("split", "imma") and "formulation" in str(e), (path, str(e))   # the model does not fit that path
            continue
        ran += 1
        assert (act == N[key + "_act"]).all() and (logits == N[key + "_logits"]).all(), (key, path)
        if h_s:
            assert (h1 == N[key + "_h1"]).all() and (c1 == N[key + "_c1"]).all(), (key, path)
    assert ran >= 2                      # automatic + the dp2a kernel, which takes any model


@pytest.mark.skipif(not RefLib.available(dropin=True), reason="oracle/_ref/libnnsp_dropin.so not built")
@pytest.mark.parametrize("nn_id", [0, 1, 2])
def test_legacy_neuralnetclass_exe_adversarial_vectors(nb, nn_id):
    """NeuralNetClass_exe (legacy symbol, incl. its debug_layer tap) of the drop-in library on the same vectors."""
    D = RefLib(dropin=True)
    key = "net_%s_acc64" % MODEL_NAME[nn_id]          # the drop-in links the model tables without -DDEF_ACC32BIT_OPT
    n0 = nb.kernel_launches()
    for i, x in enumerate(G["net_x"]):
        act, logits, h, c = D.net_eval(nn_id, x, G[key + "_h0"][i], G[key + "_c0"][i])
        assert (act == G[key + "_act"][i]).all() and (logits == G[key + "_logits"][i]).all(), (key, i)
        assert (h == G[key + "_h1"][i]).all() and (c == G[key + "_c1"][i]).all(), (key, i)
    assert nb.kernel_launches() > n0
