"""The oracle restatement against fixtures generated from the UNMODIFIED reference (tests/golden/make_golden.py).
Runs anywhere (no reference sources, no GPU)."""
import numpy as np
import pytest

from common import (GOLDEN_NETS, MODEL_NAME, NET_CASES, NET_CRAFTED, TAP_NAMES, golden, net_case_blob, net_case_dims,
                    res4, sha)

G = golden()


def test_front_end_stages_match_reference_fixture(oracle):
    wins = G["fe_windows"]
    for i, w in enumerate(wins):
        st = oracle.feature_stages(w)
        for k in ("spec", "pspec", "mel", "logmel"):
            assert (st[k] == G["fe_" + k][i]).all(), (k, i)


@pytest.mark.parametrize("acc32", [False, True])
@pytest.mark.parametrize("nn_id", [0, 1, 2])
def test_reference_wav_excerpts(oracle, nn_id, acc32):
    m = oracle.model(nn_id, acc32)
    tag = "acc32" if acc32 else "acc64"
    for wname in ("speech", "galaxy", "galaxy_s2i"):
        key = "%s_%s_%s" % (MODEL_NAME[nn_id], tag, wname)
        res, tp = oracle.nnsp_run(m, G["wav_" + wname])
        assert (res4(res) == G[key + "_res"]).all(), key
        assert [sha(getattr(tp, n)) for n in TAP_NAMES] == list(G[key + "_sha"]), key
        if not acc32:
            assert (tp.feat == G[key + "_feat"]).all() and (tp.logits == G[key + "_logits"]).all()
            assert (tp.post == G[key + "_post"]).all()


@pytest.mark.parametrize("acc32", [False, True])
@pytest.mark.parametrize("nn_id", [0, 1, 2])
def test_synthetic_streams(oracle, nb, nn_id, acc32):
    S, T = 12, 200
    x = nb.synth_pcm(S, T, first_stream=0)
    assert sha(x) == str(G["synth_pcm_sha"][0]), "synth_pcm changed: regenerate the fixtures"
    m = oracle.model(nn_id, acc32)
    key = "synth_%s_%s" % (MODEL_NAME[nn_id], "acc32" if acc32 else "acc64")
    for s in range(S):
        res, tp = oracle.nnsp_run(m, x[s])
        assert (res4(res) == G[key + "_res"][s]).all(), (key, s)
        assert [sha(getattr(tp, n)) for n in TAP_NAMES] == list(G[key + "_sha"][s]), (key, s)


@pytest.mark.parametrize("acc32", [False, True])
@pytest.mark.parametrize("nn_id", [0, 1, 2])
def test_network_corner_cases(oracle, nn_id, acc32):
    """Full-scale inputs and random/extreme LSTM state: saturation (acc64) and wrap-around (acc32)."""
    m = oracle.model(nn_id, acc32)
    key = "net_%s_%s" % (MODEL_NAME[nn_id], "acc32" if acc32 else "acc64")
    xs = G["net_x"]
    for i in range(len(xs)):
        act, logits, h, c = oracle.net_eval(m, xs[i], G[key + "_h0"][i], G[key + "_c0"][i])
        assert (act == G[key + "_act"][i]).all() and (logits == G[key + "_logits"][i]).all(), (key, i)
        assert (h == G[key + "_h1"][i]).all() and (c == G[key + "_c1"][i]).all(), (key, i)
    if nn_id == 2:
        # KWS layer 0 has qbit_kernel + qbit_input = 14 < 15: the dead accumulator-alignment shift (quirk Q1,
        # affine.c:186-187) is what these vectors pin
        assert True


def test_acc32_differs_from_acc64_somewhere():
    """The corner-case vectors really separate the two accumulator semantics."""
    diff = 0
    for name in ("s2i", "vad", "kws"):
        diff += int((G["net_%s_acc64_logits" % name] != G["net_%s_acc32_logits" % name]).any())
    assert diff > 0


def test_cascade(oracle, nb):
    S, T = 10, 2400
    x = nb.synth_pcm(S, T, first_stream=18)
    assert sha(x) == str(G["casc_pcm_sha"][0])
    models = [oracle.model(i, False) for i in range(3)]
    for s in range(S):
        res, tp, valid = oracle.cascade_run(models, x[s])
        assert (np.frombuffer(res.tobytes(), np.uint8).reshape(T, 12) == G["casc_res"][s]).all(), s
        assert (valid == G["casc_valid"][s]).all()
        assert [sha(tp.feat), sha(tp.c), sha(tp.post), sha(tp.logmel)] == list(G["casc_sha"][s])
    stages = np.unique(G["casc_res"][:, :, 0])
    assert set(stages.tolist()) == {0, 1, 2}


@pytest.mark.parametrize("acc32", [False, True])
@pytest.mark.parametrize("case", NET_CASES + NET_CRAFTED, ids=[c[0] for c in NET_CASES + NET_CRAFTED])
def test_synthetic_and_crafted_networks(oracle, case, acc32):
    """NeuralNetClass_exe of the unmodified reference on synthetic / crafted layer stacks (tests/golden/make_golden_nets.py):
    Q-format spreads, odd widths, two LSTMs, shift_32b saturation, bias << 30 wrap, tanh_fix past 5.0, cell sat32."""
    import hashlib
    N = np.load(GOLDEN_NETS)
    blob = net_case_blob(case, acc32)
    key = "%s_%s" % (case[0], "acc32" if acc32 else "acc64")
    assert hashlib.sha256(blob).hexdigest() == str(N[key + "_blob_sha"][0]), "make_blob changed: regenerate the fixture"
    a_s, h_s, n_o = net_case_dims(case)
    m = oracle.load_model(blob, acc32)
    for i, x in enumerate(N["x"]):
        act, logits, h, c = oracle.net_eval(m, x, N[key + "_h0"][i, :max(h_s, 1)], N[key + "_c0"][i, :max(h_s, 1)])
        assert (act == N[key + "_act"][i]).all() and (logits == N[key + "_logits"][i]).all(), (key, i)
        if h_s:
            assert (h[:h_s] == N[key + "_h1"][i]).all() and (c[:h_s] == N[key + "_c1"][i]).all(), (key, i)
