"""Pin the oracle: restatement (oracle/nnsp_oracle.c) == unmodified reference (oracle/_ref) on the reference's
own test wavs (where the reference tree exists), on synthetic streams and on adversarial input."""
import os

import numpy as np
import pytest

from common import TAP_NAMES, have_reference_tree
from oracle.pyoracle import RefLib, RefS2ICtrl

pytestmark = pytest.mark.skipif(not RefLib.available(False), reason="oracle/_ref not built (needs /root/reference)")
WAVS = "/root/reference/python/test_wavs"


def _same(tp_a, tp_b):
    return [n for n in TAP_NAMES if not (getattr(tp_a, n) == getattr(tp_b, n)).all()]


@pytest.mark.skipif(not have_reference_tree(), reason="reference wavs not on this machine")
@pytest.mark.parametrize("acc32", [False, True])
def test_full_reference_wavs_all_models(oracle, acc32):
    R = RefLib(acc32)
    for wname in ("speech", "galaxy", "galaxy_s2i"):
        w = np.fromfile(os.path.join(WAVS, wname + ".wav"), dtype=np.int16, offset=44)
        for nn_id in (0, 1, 2):
            r1, t1 = oracle.nnsp_run(oracle.model(nn_id, acc32), w)
            r2, t2 = R.nnsp_run(nn_id, w)
            assert (r1 == r2).all() and not _same(t1, t2), (wname, nn_id)


@pytest.mark.skipif(not have_reference_tree(), reason="reference wavs not on this machine")
def test_cascade_over_the_three_wavs(oracle):
    """config: nnCntrlClass {vad,kws,s2i} over speech+galaxy+galaxy_s2i; SURVEY.md 3.2 probe: 496/1562/942 frames."""
    w = np.concatenate([np.fromfile(os.path.join(WAVS, n + ".wav"), dtype=np.int16, offset=44)
                        for n in ("speech", "galaxy", "galaxy_s2i")])
    r1, t1, v1 = oracle.cascade_run([oracle.model(i) for i in range(3)], w)
    r2, t2, v2 = RefLib(False).cascade_run(w)
    assert (r1 == r2).all() and (v1 == v2).all()
    for n in ("logmel", "feat", "h", "c", "post"):
        assert (getattr(t1, n) == getattr(t2, n)).all(), n
    assert np.bincount(r1["stage_id"], minlength=3).tolist() == [942, 496, 1562]


@pytest.mark.parametrize("acc32", [False, True])
def test_synthetic_edge_streams(oracle, nb, acc32):
    """noise at full scale, digital silence, square wave, DC: classes 0..3 of synth_pcm, plus speech-like ones."""
    R = RefLib(acc32)
    x = nb.synth_pcm(8, 150, first_stream=16)
    for nn_id in (0, 1, 2):
        for s in range(len(x)):
            r1, t1 = oracle.nnsp_run(oracle.model(nn_id, acc32), x[s])
            r2, t2 = R.nnsp_run(nn_id, x[s])
            assert (r1 == r2).all() and not _same(t1, t2), (nn_id, s)


def test_front_end_adversarial_windows(oracle, nb):
    R = RefLib(False)
    for w in nb.adversarial_windows():
        a, b = oracle.feature_stages(w), R.feature_stages(w)
        for k in a:
            assert (a[k] == b[k]).all(), k


@pytest.mark.parametrize("acc32", [False, True])
def test_network_random_state_and_full_scale_input(oracle, acc32):
    R = RefLib(acc32)
    rng = np.random.default_rng(5)
    for nn_id in (0, 1, 2):
        _, h_s, _ = R.strides(nn_id)
        m = oracle.model(nn_id, acc32)
        for k in range(12):
            x = rng.integers(-32768, 32768, 240).astype(np.int16) if k else np.full(240, -32768, np.int16)
            h = rng.integers(-32768, 32768, h_s).astype(np.int16)
            c = rng.integers(-2 ** 31, 2 ** 31, h_s).astype(np.int32)
            a = oracle.net_eval(m, x, h, c)
            b = R.net_eval(nn_id, x, h, c)
            for u, v in zip(a, b):
                assert (u == v).all(), (nn_id, k)


def test_reset_of_a_live_instance(oracle, nb):
    R = RefLib(False)
    x = nb.synth_pcm(1, 100, first_stream=5)[0]
    m = oracle.model(1, False)
    st = oracle.lib.nnsp_oracle_stream_new()
    oracle.nnsp_run(m, x[:50 * 160], state=st, reset=1, taps=False)
    r1, t1 = oracle.nnsp_run(m, x[50 * 160:], state=st, reset=2)
    oracle.lib.nnsp_oracle_stream_free(st)
    R.nnsp_run(1, x[:50 * 160], reset=1, taps=False)
    r2, t2 = R.nnsp_run(1, x[50 * 160:], reset=2)
    assert (r1 == r2).all() and not _same(t1, t2)


@pytest.mark.skipif(not RefS2ICtrl.available(), reason="oracle/_ref/libnnsp_ref_s2ictrl.so not built")
def test_s2i_only_controller_is_the_single_stage_cascade(oracle, nb):
    """evb/src/s2iCntrlClass.c:92-122 (compiled unmodified) == nnCntrlClass semantics with seq = {s2i}, look-back 0:
    the equivalence INTEGRATION.md claims, pinned frame by frame (detections, outputs, resets on detection)."""
    R = RefS2ICtrl()
    x = np.concatenate([nb.synth_pcm(12, 700, first_stream=17), nb.synth_pcm(1, 700, first_stream=5)])   # 27, 28 and 5 detect
    par = oracle.default_params()
    par[2] = 0                       # frs_vbufBk_s2i: the S2I-only controller reads the newest frame (s2iCntrlClass.c:105-109)
    om = [oracle.model(i) for i in range(3)]
    detections = 0
    for th_cnt in (4, 1):
        par[5] = th_cnt              # thresh_cnts_s2i
        for s in range(len(x)):
            want = R.run(x[s], reset=1, thresh_cnts=th_cnt)
            got, _, _ = oracle.cascade_run(om, x[s], seq=(0,), params=par, taps=False)
            for f in ("stage_id", "pos_after", "detected", "outputs"):
                assert (got[f] == want[f]).all(), (s, f)
            detections += int(want["detected"].sum())
    assert detections > 0            # the controller really reset instances on the way


def test_controller_state_swap_keeps_streams_apart(nb):
    """ref_cascade_batch (the CPU arm of bench.py): streams served in turn by the one reference controller, state swapped
    in and out per chunk, give exactly what each stream gives alone in one go."""
    R = RefLib(False)
    S, T = 5, 260
    x = nb.synth_pcm(S, T, first_stream=18)
    states = np.zeros((S, R.cascade_state_bytes()), np.uint8)
    parts = [R.cascade_batch(x[:, :100 * 160], states, fresh=True, want_results=True),
             R.cascade_batch(x[:, 100 * 160:101 * 160], states, fresh=False, want_results=True),
             R.cascade_batch(x[:, 101 * 160:], states, fresh=False, want_results=True)]
    got = np.concatenate(parts, axis=1)
    for s in range(S):
        want, _, _ = R.cascade_run(x[s], taps=False)
        assert (got[s] == want).all(), s


@pytest.mark.parametrize("acc32", [False, True])
def test_synthetic_network_through_the_reference(oracle, acc32):
    """restatement == unmodified reference on a synthetic two-LSTM stack, fresh random inputs (not the fixture's)."""
    from common import NET_CASES, NET_CRAFTED, net_case_blob, net_case_dims
    R = RefLib(False)
    rng = np.random.default_rng(99)
    for case in (NET_CASES[1], NET_CRAFTED[0], NET_CRAFTED[2]):
        blob = net_case_blob(case, acc32)
        a_s, h_s, n_o = net_case_dims(case)
        m = oracle.load_model(blob, acc32)
        for k in range(6):
            x = rng.integers(-32768, 32768, 240).astype(np.int16)
            h = rng.integers(-32768, 32768, h_s).astype(np.int16)
            c = rng.integers(-2 ** 31, 2 ** 31, h_s).astype(np.int32)
            a = oracle.net_eval(m, x, h, c)
            b = R.net_eval_blob(blob, acc32, x, h, c, a_s, n_o)
            for u, v in zip(a, b):
                assert (u[: len(v)] == v).all(), (case[0], k)


def test_params_rewritten_on_a_live_controller(oracle, nb):
    """ParamCntrlClass is a member of the controller instance and the NNSPClass instances read the thresholds through
    pointers into it (nnCntrlClass.c:100-123), the time-outs are read directly (:185, :219): an application that rewrites
    them between frames is obeyed from the next frame on -- including a time-out shortened below the running counter.
    The restatement must do the same as the reference itself."""
    R = RefLib(False)
    om = [oracle.model(i) for i in range(3)]
    base = oracle.default_params()
    names = ["thresh_prob_vad", "thresh_cnts_vad", "frs_vbufBk_s2i", "thresh_timeout_s2i", "thresh_prob_s2i", "thresh_cnts_s2i",
             "frs_vbufBk_kws", "thresh_timeout_kws", "thresh_prob_kws", "thresh_cnts_kws"]
    for first in (15, 5, 47):                                    # streams that sit in KWS / S2I when the parameters change
        x = nb.synth_pcm(1, 700, first_stream=first)[0]
        p1, p2 = base.copy(), base.copy()
        p1[names.index("thresh_timeout_kws")] = 400; p1[names.index("thresh_timeout_s2i")] = 300
        p2[names.index("thresh_timeout_kws")] = 31; p2[names.index("thresh_timeout_s2i")] = 23       # below the counters of the first part
        p2[names.index("thresh_cnts_vad")] = 2; p2[names.index("thresh_prob_kws")] = 900
        st = oracle.lib.nnsp_oracle_cascade_new()
        a1, _, _ = oracle.cascade_run(om, x[:300 * 160], params=p1, state=st, reset=1, taps=False)
        a2, _, _ = oracle.cascade_run(om, x[300 * 160:], params=p2, state=st, reset=0, taps=False)
        oracle.lib.nnsp_oracle_cascade_free(st)
        b1, _, _ = R.cascade_run(x[:300 * 160], params=p1, reset=1, taps=False)
        b2, _, _ = R.cascade_run(x[300 * 160:], params=p2, reset=0, taps=False)
        assert (a1 == b1).all() and (a2 == b2).all(), first
        keep, _, _ = R.cascade_run(np.concatenate([x[:300 * 160], x[300 * 160:]]), params=p1, reset=1, taps=False)
        assert not (keep[300:] == b2).all(), "stream %d: the rewritten parameters changed nothing, pick another stream" % first
