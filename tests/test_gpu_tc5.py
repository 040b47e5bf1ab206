"""GPU parity of the tcgen05 layer-0 kernel (nnsp_b200/csrc/nnsp_tc5.cuh) that the batched scan-split path runs when no
activation tap is requested: first fc layer of NeuralNetClass_exe (neural_nets.c:44-168, affine.c:409-490, activation.c:31-69)
for all (stream, inference) rows of a call on the 5th-generation tensor cores. Checked against the oracle through what the
kernel feeds: the result records of the untapped calls, and every tap of a final tapped call whose LSTM state and context
were produced by the tcgen05 calls before it."""
import numpy as np
import pytest

from common import NET_CASES, make_blob

pytestmark = pytest.mark.gpu
TAPS = ["feat", "act", "logits", "hstate", "cstate", "post"]
ORACLE_TAP = dict(feat="feat", act="act", logits="logits", hstate="h", cstate="c", post="post")
MODEL_FILE = {0: "s2i.nnspm", 1: "vad.nnspm", 2: "kws_galaxy.nnspm"}
# call lengths: one-frame calls (both phases of the stride-2 gate), a call with a single inference, 100 frames (the bench
# shape), 129 and 257 frames (65 / 129 inferences: two and three 62-inference chunks per stream pair)
LENS = [1, 1, 2, 7, 100, 129, 3, 257, 64]
LAST = 9


@pytest.fixture(autouse=True)
def _force_tc5(monkeypatch):
    """the library takes the kernel on its own only when the tiles are at least 45 % full (calls of >= 58 frames); the
    small shapes below are forced through it"""
    monkeypatch.setenv("NNSP_B200_TC5", "2")


def _run(nb, oracle, m, m_or, S, first_stream, h_stride=None):
    T = sum(LENS) + LAST
    pcm = nb.synth_pcm(S, T, first_stream=first_stream)
    before = nb.tc5_launches()
    b = nb.NNSPBatch(m, S)
    assert b.nn_path == "split"
    parts, t = [], 0
    for n in LENS:
        parts.append(b.exec(pcm[:, t * 160:(t + n) * 160]))
        t += n
    ran = nb.tc5_launches() - before
    assert len(LENS) - 2 <= ran <= len(LENS), "the tcgen05 kernel did not run for every untapped call with an inference"   # a one-frame call may hold none
    res_last, taps = b.exec(pcm[:, t * 160:], taps=True)            # taps: the mma.sync kernel, on the state the calls above left
    assert nb.tc5_launches() - before == ran
    b.close()
    got = np.concatenate(parts + [res_last], axis=1)
    kw = {} if h_stride is None else dict(h_stride=h_stride)
    for s in range(S):
        r, tp = oracle.nnsp_run(m_or, pcm[s], **kw)
        assert (r == got[s]).all(), "results differ on stream %d: first frame %d" % (s, int(np.nonzero(r != got[s])[0][0]))
        for name in TAPS:
            a, o = taps[name][s], getattr(tp, ORACLE_TAP[name])[t:]
            assert a.shape == o.shape and (a == o).all(), "tap %s of the last call differs on stream %d" % (name, s)


@pytest.mark.parametrize("S", [1, 2, 3, 17, 35])
@pytest.mark.parametrize("nn_id,acc32", [(0, False), (1, False), (2, True)])
def test_shipped_models_untapped_calls_match_oracle(nb, oracle, nn_id, acc32, S):
    m = nb.Model.from_blob(nb.MODEL_DIR + "/" + MODEL_FILE[nn_id], acc32=acc32)
    _run(nb, oracle, m, oracle.model(nn_id, acc32), S, first_stream=10 * nn_id + S)


ELIGIBLE = [c for c in NET_CASES if c[0] in ("two_lstm", "odd_widths", "big_shifts")]     # tanh layer 0 with the exact 32-bit finish, LSTM behind it


@pytest.mark.parametrize("case", ELIGIBLE, ids=[c[0] for c in ELIGIBLE])
def test_synthetic_widths_and_shifts(nb, oracle, case):
    """units 24 (padded to 32), 7 (padded to 16: 9 zero-weight units), 16 with a finish shift far from the shipped ones"""
    name, nn_id, sizes, types, acts, qk, qi, qb = case
    raw = make_blob(nn_id, sizes, types, list(acts), list(qk), list(qi), list(qb), seed=len(name))
    h_stride = sum(sizes[i + 1] for i, t in enumerate(types) if t == 1) or 1
    _run(nb, oracle, nb.Model.from_blob(raw), oracle.load_model(raw, False), 21, first_stream=400, h_stride=h_stride)


def test_ineligible_layers_keep_the_mma_sync_kernel(nb, oracle):
    """relu6 / sigmoid first layers and stacks without an LSTM behind layer 0 are not taken by the tcgen05 kernel"""
    for case in NET_CASES:
        name, nn_id, sizes, types, acts, qk, qi, qb = case
        if name not in ("fc_only", "lstm_wide", "lstm_last_fc_sigmoid"):
            continue
        raw = make_blob(nn_id, sizes, types, list(acts), list(qk), list(qi), list(qb), seed=3)
        m = nb.Model.from_blob(raw)
        b = nb.NNSPBatch(m, 9)
        before = nb.tc5_launches()
        pcm = nb.synth_pcm(9, 30, first_stream=5)
        got = b.exec(pcm)
        assert nb.tc5_launches() == before, name
        m_or = oracle.load_model(raw, False)
        h_stride = sum(sizes[i + 1] for i, t in enumerate(types) if t == 1) or 1
        for s in range(9):
            r, _ = oracle.nnsp_run(m_or, pcm[s], h_stride=h_stride, taps=False)
            assert (r == got[s]).all(), (name, s)
        b.close()


def test_short_calls_keep_the_mma_sync_kernel(nb, monkeypatch):
    """the selection rule: tiles of 2 streams x 64 inference slots must be at least 45 % in use"""
    monkeypatch.delenv("NNSP_B200_TC5")
    m = nb.Model.from_blob(nb.MODEL_DIR + "/" + MODEL_FILE[1])
    b = nb.NNSPBatch(m, 8)
    for frames, want in [(2, 0), (16, 0), (40, 0), (64, 1), (100, 1), (130, 1), (300, 1)]:
        before = nb.tc5_launches()
        b.exec(nb.synth_pcm(8, frames, first_stream=1))
        assert nb.tc5_launches() - before == want, frames
    b.close()


def test_async_host_calls_of_bench_shape(nb, oracle, monkeypatch):
    """the end-to-end loop of bench.py at 100 frames per call: the kernel is chosen by the library's own rule and runs on
    the pipeline slices of asynchronous host-buffer calls (600 streams = 2 slices; slice starts are multiples of 16)"""
    monkeypatch.delenv("NNSP_B200_TC5")
    S, n, calls = 600, 100, 3
    pcm = nb.synth_pcm(S, n * calls, first_stream=900)
    m = nb.Model.from_blob(nb.MODEL_DIR + "/" + MODEL_FILE[2], acc32=True)
    b = nb.NNSPBatch(m, S)
    pin = [nb.PinnedArray((S, n * 160), np.int16) for _ in range(2)]
    pres = [nb.PinnedArray((S, n), nb.RESULT_DT) for _ in range(2)]
    before, got, prev = nb.tc5_launches(), [], None
    for k in range(calls):
        pin[k & 1].array[...] = pcm[:, k * n * 160:(k + 1) * n * 160]
        tk = b.exec_host_async(pin[k & 1].array, pres[k & 1].array)
        if prev is not None:
            b.wait_host(prev)
            got.append(pres[(k - 1) & 1].array.copy())
        prev = tk
    b.wait_host(prev)
    got.append(pres[(calls - 1) & 1].array.copy())
    assert nb.tc5_launches() - before >= calls
    got = np.concatenate(got, axis=1)
    m_or = oracle.model(2, True)
    for s in list(range(0, S, 41)) + [S - 1]:
        r, _ = oracle.nnsp_run(m_or, pcm[s], taps=False)
        assert (r == got[s]).all(), s
    for x in pin + pres:
        x.free()
    b.close()


def test_cascade_first_round_on_the_tcgen05_kernel(nb, oracle):
    """NNSP_B200_TC5=2 (the fixture) also sends layer 0 of the cascade's first round through vseq_kernel + the tcgen05 kernel
    (device-side stream lists, per-stream first inference frames, context / look-back / fresh-instance rows): every field of
    the result records against the oracle, over calls long enough for the kernel to be chosen and with stage changes inside."""
    S, n, calls = 48, 150, 5
    pcm = nb.synth_pcm(S, n * calls, first_stream=0)
    models = [nb.Model.from_blob(nb.MODEL_DIR + "/" + MODEL_FILE[i]) for i in range(3)]
    c = nb.Cascade(models, S, params=dict(thresh_timeout_s2i=45))
    before = nb.tc5_launches()
    got = np.concatenate([c.exec(pcm[:, k * n * 160:(k + 1) * n * 160]) for k in range(calls)], axis=1)
    assert nb.tc5_launches() - before == 3 * calls          # one per model group of the first round
    par = c.params_array()
    c.close()
    om = [oracle.model(i) for i in range(3)]
    stages = np.zeros(3, np.int64)
    for s in range(S):
        r, _, _ = oracle.cascade_run(om, pcm[s], params=par, taps=False)
        for f in r.dtype.names:
            assert (got[s][f] == r[f]).all(), "stream %d field %s" % (s, f)
        stages += np.bincount(r["stage_id"], minlength=3)
    assert (stages > 0).all(), stages                        # all three models were live somewhere
