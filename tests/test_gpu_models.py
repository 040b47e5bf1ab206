"""GPU parity on SYNTHETIC models: layer stacks the shipped tables never exercise (no LSTM, two LSTMs, widths that are
not multiples of 8 / 32, wide Q-format spreads, ACC32BIT_OPT wrap) through every network path that accepts them,
against the oracle on every tap. The default (auto) path picks scan-split when every layer has the exact 32-bit finish,
else IMMA in the time loop, else dp2a -- so the fallbacks are exercised too."""
import numpy as np
import pytest

from common import NET_CASES, make_blob

pytestmark = pytest.mark.gpu
TAPS = ["feat", "act", "logits", "hstate", "cstate", "post"]
ORACLE_TAP = dict(feat="feat", act="act", logits="logits", hstate="h", cstate="c", post="post")

CASES = NET_CASES


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
