"""The CUDA engine against the fixtures generated from the unmodified reference (no oracle in the loop)."""
import numpy as np
import pytest

from common import MODEL_NAME, golden, res4, sha

pytestmark = pytest.mark.gpu
G = golden()
MODEL_FILE = {0: "s2i.nnspm", 1: "vad.nnspm", 2: "kws_galaxy.nnspm"}
GPU_TAPS = ["logmel", "feat", "act", "logits", "hstate", "cstate", "post"]


def test_front_end_fixture(nb):
    out = nb.feature_stages(G["fe_windows"])
    for k in ("spec", "pspec", "mel", "logmel"):
        assert (out[k] == G["fe_" + k]).all(), k


@pytest.mark.parametrize("acc32", [False, True])
@pytest.mark.parametrize("nn_id", [0, 1, 2])
def test_reference_wav_excerpts_and_synthetic_streams(nb, nn_id, acc32):
    m = nb.Model.from_blob(nb.MODEL_DIR + "/" + MODEL_FILE[nn_id], acc32=acc32)
    tag = "acc32" if acc32 else "acc64"
    wavs = np.stack([G["wav_" + w] for w in ("speech", "galaxy", "galaxy_s2i")])
    b = nb.NNSPBatch(m, 3)
    res, taps = b.exec(wavs, taps=True)
    for i, w in enumerate(("speech", "galaxy", "galaxy_s2i")):
        key = "%s_%s_%s" % (MODEL_NAME[nn_id], tag, w)
        assert (res4(res[i]) == G[key + "_res"]).all(), key
        assert [sha(taps[n][i]) for n in GPU_TAPS] == list(G[key + "_sha"]), key
    b.close()
    S, T = 12, 200
    x = nb.synth_pcm(S, T)
    b = nb.NNSPBatch(m, S)
    res, taps = b.exec(x, taps=True)
    key = "synth_%s_%s" % (MODEL_NAME[nn_id], tag)
    for s in range(S):
        assert (res4(res[s]) == G[key + "_res"][s]).all(), (key, s)
        assert [sha(taps[n][s]) for n in GPU_TAPS] == list(G[key + "_sha"][s]), (key, s)
    b.close()


def test_cascade_fixture(nb):
    S, T = 10, 2400
    x = nb.synth_pcm(S, T, first_stream=18)
    c = nb.Cascade([nb.Model.from_blob(nb.MODEL_DIR + "/" + MODEL_FILE[i]) for i in range(3)], S)
    res, taps = c.exec(x, taps=True)
    for s in range(S):
        assert (np.frombuffer(res[s].tobytes(), np.uint8).reshape(T, 12) == G["casc_res"][s]).all(), s
        assert [sha(taps["feat"][s]), sha(taps["cstate"][s]), sha(taps["post"][s]), sha(taps["logmel"][s])] == list(G["casc_sha"][s]), s
    c.close()
