"""The audio front door (nnsp_b200_wav_*): RIFF/WAVE files -> [stream][frame][160] int16 PCM. CPU only (host code).
Files are written here with the standard library and, for the odd cases, byte by byte."""
import os
import struct
import wave

import numpy as np
import pytest

from common import golden


def _write(path, x, rate=16000, channels=1):
    with wave.open(str(path), "wb") as w:
        w.setnchannels(channels)
        w.setsampwidth(2)
        w.setframerate(rate)
        w.writeframes(np.ascontiguousarray(x, "<i2").tobytes())


def test_reads_back_what_the_standard_library_wrote(nb, tmp_path):
    rng = np.random.default_rng(3)
    lens = [160 * 7, 160 * 7 + 59, 0, 160 * 3]
    xs = [rng.integers(-32768, 32768, n).astype(np.int16) for n in lens]
    paths = []
    for i, x in enumerate(xs):
        p = tmp_path / ("s%d.wav" % i)
        _write(p, x)
        paths.append(str(p))
        assert nb.wav_info(p) == (16000, 1, 16, len(x))
    pcm, got = nb.wav_load_streams(paths, 8)
    assert got.tolist() == [7, 8, 0, 3]
    for i, x in enumerate(xs):
        want = np.zeros(8 * 160, np.int16)
        want[: len(x)] = x[: 8 * 160]
        assert (pcm[i] == want).all(), i
    pcm2, got2 = nb.wav_load_streams(paths[:2], 3, first_frame=6)               # a window that runs past the end
    assert got2.tolist() == [1, 2] and (pcm2[1][:160 + 59] == xs[1][6 * 160:]).all() and (pcm2[1][160 + 59:] == 0).all()


def test_channel_pick_extensible_header_and_extra_chunks(nb, tmp_path):
    rng = np.random.default_rng(4)
    x = rng.integers(-32768, 32768, (800, 3)).astype("<i2")                      # 3 interleaved channels
    p = tmp_path / "multi.wav"
    _write(p, x, channels=3)
    for ch in range(3):
        pcm, got = nb.wav_load_streams([str(p)], 5, channel=ch)
        assert got[0] == 5 and (pcm[0] == x[:, ch]).all()
    # WAVE_FORMAT_EXTENSIBLE fmt chunk (40 bytes), a LIST chunk of odd length in front of the data chunk
    data = x[:, 0].tobytes()
    fmt = struct.pack("<HHIIHHHHIH14s", 0xFFFE, 1, 16000, 32000, 2, 16, 22, 16, 4, 1, b"\x00\x00\x00\x00\x10\x00\x80\x00\x00\xaa\x00\x38\x9b\x71")
    body = b"WAVE" + b"LIST" + struct.pack("<I", 5) + b"abcde\x00" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"data" + struct.pack("<I", len(data)) + data
    q = tmp_path / "ext.wav"
    q.write_bytes(b"RIFF" + struct.pack("<I", len(body)) + body)
    assert nb.wav_info(q) == (16000, 1, 16, 800)
    pcm, _ = nb.wav_load_streams([str(q)], 5)
    assert (pcm[0] == x[:, 0]).all()


def test_what_is_refused(nb, tmp_path):
    x = np.zeros(1600, np.int16)
    p = tmp_path / "r8k.wav"
    _write(p, x, rate=8000)
    with pytest.raises(nb.NnspError, match="no resampler"):
        nb.wav_load_streams([str(p)], 2)
    q = tmp_path / "notwav.bin"
    q.write_bytes(b"hello, this is not audio")
    with pytest.raises(nb.NnspError, match="RIFF"):
        nb.wav_info(q)
    with pytest.raises(nb.NnspError, match="cannot open"):
        nb.wav_info(tmp_path / "missing.wav")
    with wave.open(str(tmp_path / "w8.wav"), "wb") as w:                          # 8-bit samples
        w.setnchannels(1); w.setsampwidth(1); w.setframerate(16000); w.writeframes(bytes(1600))
    with pytest.raises(nb.NnspError, match="16-bit"):
        nb.wav_load_streams([str(tmp_path / "w8.wav")], 2)
    _write(tmp_path / "ok.wav", x)
    with pytest.raises(nb.NnspError, match="channel"):
        nb.wav_load_streams([str(tmp_path / "ok.wav")], 2, channel=1)


@pytest.mark.gpu
def test_wav_files_end_to_end_against_the_reference_fixture(nb, tmp_path):
    """The reference's own test clips (2.5 s excerpts of python/test_wavs/*.wav, stored in the golden fixture), written as
    WAVE files, loaded through the front door and run through the host-buffer entry point: results equal what the
    unmodified reference produced on those clips (tests/golden/make_golden.py)."""
    G = golden()
    names = ("speech", "galaxy", "galaxy_s2i")
    paths = []
    for n in names:
        _write(tmp_path / (n + ".wav"), G["wav_" + n])
        paths.append(str(tmp_path / (n + ".wav")))
    T = len(G["wav_speech"]) // 160
    pcm, got = nb.wav_load_streams(paths, T)
    assert got.tolist() == [T] * 3
    for nn_id, mname, f in ((0, "s2i", "s2i.nnspm"), (1, "vad", "vad.nnspm"), (2, "kws", "kws_galaxy.nnspm")):
        b = nb.NNSPBatch(nb.Model.from_blob(os.path.join(nb.MODEL_DIR, f)), 3)
        res = b.exec_host(pcm)
        b.close()
        for i, n in enumerate(names):
            assert (res[i].view(np.int16).reshape(T, 4) == G["%s_acc64_%s_res" % (mname, n)]).all(), (mname, n)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["batch", "cascade"])
def test_audadc_words_conditioned_on_the_device_in_front_of_the_path(nb, oracle, kind):
    """NNSP_B200_HOST_AUDADC: raw 32-bit AUDADC words over the link, audio_frame_callback's conditioning
    (main_nnsp.cc:58-65) on the device, then the path -- against the oracle fed with the CPU-conditioned PCM."""
    S, T = 300, 64
    rng = np.random.default_rng(8)
    pcm = nb.synth_pcm(S, T, first_stream=40)
    raw = (pcm.astype(np.int64) & 0xffff).astype(np.uint32) | (rng.integers(0, 1 << 16, pcm.shape).astype(np.uint32) << 16)
    raw ^= rng.integers(0, 16, pcm.shape).astype(np.uint32)                       # junk in the 4 bits below the sample
    want_pcm = oracle.ingest_audadc(raw)
    if kind == "batch":
        h = nb.NNSPBatch(nb.Model.from_blob(os.path.join(nb.MODEL_DIR, "vad.nnspm")), S)
        ref = nb.NNSPBatch(nb.Model.from_blob(os.path.join(nb.MODEL_DIR, "vad.nnspm")), S)
    else:
        models = [nb.Model.from_blob(os.path.join(nb.MODEL_DIR, f)) for f in ("s2i.nnspm", "vad.nnspm", "kws_galaxy.nnspm")]
        h, ref = nb.Cascade(models, S), nb.Cascade(models, S)
    h.set_host_format("audadc")
    got = np.concatenate([h.exec_host(np.ascontiguousarray(raw[:, :24 * 160])), h.exec_host(np.ascontiguousarray(raw[:, 24 * 160:]))], axis=1)
    want = ref.exec_host(want_pcm)
    for f in want.dtype.names:
        assert (got[f] == want[f]).all(), f
    if kind == "batch":                                                           # and the oracle itself on a few streams
        m = oracle.model(1, False)
        for s in range(0, S, 29):
            r, _ = oracle.nnsp_run(m, want_pcm[s], taps=False)
            assert (r == got[s]).all(), s
    h.close(); ref.close()
