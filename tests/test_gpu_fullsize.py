"""BASELINE.json's full-size configurations on the GPU, checked through size-independent properties:
  * position independence -- a stream's results depend only on its own PCM, so every copy of a base stream, wherever it
    sits among tens of thousands (any tile, any slice, any group of the cascade), must produce identical records;
  * a sample of the base streams against the CPU oracle, every frame.
cfg 2: VAD x 4 096; cfg 3: KWS x 16 384 with ACC32BIT_OPT semantics; cfg 4: S2I x 32 768; cfg 5: the cascade at its
per-GPU share (65 536 / 8 = 8 192 streams)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
MODEL_FILE = {0: "s2i.nnspm", 1: "vad.nnspm", 2: "kws_galaxy.nnspm"}
BASE, T = 192, 100


def _pcm(nb, S):
    base = nb.synth_pcm(BASE, T, first_stream=11)
    reps = (S + BASE - 1) // BASE
    idx = (np.arange(S) * 7 + (np.arange(S) // BASE) * 13) % BASE      # copies land on different tile rows / slices
    return base, idx, np.ascontiguousarray(base[idx])


@pytest.mark.parametrize("nn_id,acc32,S", [(1, False, 4096), (2, True, 16384), (0, False, 32768)])
def test_full_size_batch(nb, oracle, nn_id, acc32, S):
    base, idx, pcm = _pcm(nb, S)
    m = nb.Model.from_blob(nb.MODEL_DIR + "/" + MODEL_FILE[nn_id], acc32=acc32)
    b = nb.NNSPBatch(m, S)
    assert b.nn_path == "split"
    res = b.exec(pcm[:, :60 * 160])
    res = np.concatenate([res, b.exec(pcm[:, 60 * 160:])], axis=1).view(np.int16).reshape(S, T, 4)
    host = b2 = None
    first = np.full(BASE, -1)
    for s in range(S):                                  # first occurrence of every base stream
        if first[idx[s]] < 0:
            first[idx[s]] = s
    assert (res == res[first[idx]]).all(), "results depend on where a stream sits in the batch"
    m_or = oracle.model(nn_id, acc32)
    for k in range(0, BASE, 6):
        r, _ = oracle.nnsp_run(m_or, base[k], taps=False)
        assert (r.view(np.int16).reshape(T, 4) == res[first[k]]).all(), "base stream %d" % k
    b.close()
    b2 = nb.NNSPBatch(m, S)                             # the host-buffer entry point slices the same work 16 ways
    host = b2.exec_host(pcm).view(np.int16).reshape(S, T, 4)
    assert (host == res).all()
    b2.close()


def test_full_size_cascade(nb, oracle):
    S = 8192
    base, idx, pcm = _pcm(nb, S)
    models = [nb.Model.from_blob(nb.MODEL_DIR + "/" + MODEL_FILE[i]) for i in range(3)]
    params = dict(frs_vbufBk_kws=80, frs_vbufBk_s2i=80, thresh_timeout_kws=120, thresh_timeout_s2i=90)   # stage changes within 300 frames
    c = nb.Cascade(models, S, params=params)
    parts = [c.exec(pcm) for _ in range(3)]             # the same second of audio three times: 300 frames per stream
    res = np.concatenate(parts, axis=1)
    first = np.full(BASE, -1)
    for s in range(S):
        if first[idx[s]] < 0:
            first[idx[s]] = s
    for f in res.dtype.names:
        assert (res[f] == res[f][first[idx]]).all(), "cascade field %s depends on the stream's position" % f
    om = [oracle.model(i, False) for i in range(3)]
    par = c.params_array()
    stages = set()
    for k in range(0, BASE, 8):
        r, _, _ = oracle.cascade_run(om, np.tile(base[k], 3), params=par, taps=False)
        for f in ("stage_id", "pos_after", "detected", "outputs", "cnt_timeout"):
            assert (res[first[k]][f] == r[f]).all(), "base stream %d field %s" % (k, f)
        stages |= set(np.unique(r["stage_id"]).tolist())
    assert len(stages) >= 2
    c.close()
