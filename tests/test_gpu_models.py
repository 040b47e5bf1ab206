"""GPU parity on SYNTHETIC models: layer stacks the shipped tables never exercise (no LSTM, two LSTMs, widths that are
not multiples of 8 / 32, wide Q-format spreads, ACC32BIT_OPT wrap) through every network path that accepts them,
against the oracle on every tap. The default (auto) path picks scan-split when every layer has the exact 32-bit finish,
else IMMA in the time loop, else dp2a -- so the fallbacks are exercised too."""
import numpy as np
import pytest

from common import make_blob

pytestmark = pytest.mark.gpu
TAPS = ["feat", "act", "logits", "hstate", "cstate", "post"]
ORACLE_TAP = dict(feat="feat", act="act", logits="logits", hstate="h", cstate="c", post="post")

CASES = [
    # nn_id, sizes, types (0 fc, 1 lstm), acts (0 relu6 1 tanh 2 sigmoid 3 linear), qk, qi, qb
    ("fc_only", 1, (240, 33, 17, 2), (0, 0, 0), (1, 0, 3), (7, 5, 6), (8, 15, 12), (14, 15, 15)),
    ("two_lstm", 2, (240, 24, 20, 12, 9, 2), (0, 1, 0, 1, 0), (1, 1, 2, 1, 3), (6, 5, 5, 5, 6), (8, 15, 15, 15, 15), (13, 13, 15, 14, 15)),
    ("lstm_wide", 0, (240, 40, 100, 41), (0, 1, 0), (0, 1, 3), (7, 4, 5), (8, 15, 15), (14, 14, 14)),        # 13 unit groups
    ("lstm_requant", 0, (240, 40, 100, 41), (0, 1, 0), (0, 1, 3), (7, 4, 5), (8, 12, 15), (14, 14, 14)),     # qbit_input_rec != qbit_input
    ("lstm_last_fc_sigmoid", 1, (240, 10, 6, 2), (0, 1, 0), (2, 1, 3), (7, 6, 7), (8, 15, 15), (15, 12, 15)),
    ("odd_widths", 2, (240, 7, 5, 3, 2), (0, 1, 0, 0), (1, 1, 0, 3), (7, 5, 5, 7), (8, 15, 15, 12), (14, 13, 15, 15)),
    ("big_shifts", 1, (240, 16, 16, 2), (0, 1, 0), (1, 1, 3), (3, 7, 2), (8, 15, 15), (6, 15, 4)),
    ("bias_shift_18", 1, (240, 16, 16, 2), (0, 1, 0), (1, 1, 3), (12, 7, 2), (8, 15, 15), (2, 15, 4)),  # layer 0 needs the 64-bit finish
]


@pytest.mark.parametrize("acc32", [False, True])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_synthetic_models_match_oracle(nb, oracle, case, acc32):
    name, nn_id, sizes, types, acts, qk, qi, qb = case
    raw = make_blob(nn_id, sizes, types, list(acts), list(qk), list(qi), list(qb), seed=len(name) + 7 * acc32)
    m = nb.Model.from_blob(raw, acc32=acc32)
    m_or = oracle.load_model(raw, acc32)
    S, T = 37, 41
    pcm = nb.synth_pcm(S, T, first_stream=50)
    h_stride = sum(sizes[i + 1] for i, t in enumerate(types) if t == 1) or 1
    paths = ["auto", "dp2a"]
    for path in paths:
        try:
            b = nb.NNSPBatch(m, S, nn_path=path)
        except nb.NnspError:
            continue
        if path == "auto":          # which kernels "automatic" means for this stack
            want = "split" if name not in ("bias_shift_18", "lstm_requant") else ("imma", "dp2a")
            assert b.nn_path == want or b.nn_path in want, (name, b.nn_path)
        res, taps = b.exec(pcm, taps=True)
        res2 = np.concatenate([b.exec(pcm[:, :13 * 160]), b.exec(pcm[:, 13 * 160:])], axis=1) if path == "auto" else None
        for s in range(S):
            r, tp = oracle.nnsp_run(m_or, pcm[s], h_stride=h_stride)
            assert (r == res[s]).all(), "%s/%s: results differ on stream %d" % (name, path, s)
            for t in TAPS:
                a, o = taps[t][s], getattr(tp, ORACLE_TAP[t])
                assert a.shape == o.shape and (a == o).all(), "%s/%s: tap %s differs on stream %d" % (name, path, t, s)
        b.close()
        if res2 is not None:
            # a second handle: first the same frames without taps (the pipelined product path), continued in two calls
            b = nb.NNSPBatch(m, S, nn_path=path)
            a1 = b.exec(pcm[:, :13 * 160]); a2 = b.exec(pcm[:, 13 * 160:])
            b.close()
            got = np.concatenate([a1, a2], axis=1)
            for s in range(S):
                r, _ = oracle.nnsp_run(m_or, pcm[s], h_stride=h_stride, taps=False)
                assert (r == got[s]).all(), "%s/%s (no taps, chunked): stream %d" % (name, path, s)
