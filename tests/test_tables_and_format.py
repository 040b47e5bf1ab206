"""Constant tables, model container and weight interleave (host logic, no GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from common import ROOT, have_reference_tree
from oracle.pyoracle import RefLib

TABLES = ["stft_win", "fft_tw", "rfft_tw", "bitrev", "mel", "log_lut", "tanh_lut"]


def test_generated_tables_have_the_pinned_fingerprint(nb):
    # nnsp_b200_table returns NULL tables (negative count) if the self check failed
    for name in TABLES:
        assert len(nb.table(name)) > 0


@pytest.mark.skipif(not RefLib.available(False), reason="oracle/_ref not built")
def test_generated_tables_equal_the_reference_objects(nb):
    """window_stft_coef.c, twiddle_fft_dif.c, melSpec_coeff.c, fixlog10.c:6, activation.c:5 -- as compiled."""
    R = RefLib(False)
    for name in TABLES:
        assert (nb.table(name) == R.table(name)).all(), name


def test_window_bound_behind_the_elided_fft_clamps(nb):
    """sum |window| * 32768 >> 15 bounds every FFT intermediate: the sat32 clamps of complex.c:49-51,67-69 and
    fft.c:112-114 cannot fire and the power spectrum stays below 2^31 (SURVEY.md hazard H5)."""
    win = nb.table("stft_win").astype(np.int64)
    bound = int(((win * 32768) >> 15).sum())
    assert bound == 8175452
    growth = (1 + 2.0 ** -14) ** 5                      # Q15 twiddles can exceed unit modulus by < 2^-14 per stage
    assert bound * growth * np.sqrt(2) < 2 ** 31 / 64   # far below the clamp
    assert (bound * growth) ** 2 / 32768 < 2 ** 31      # |X|^2 >> 15 fits a positive int32


def test_blob_round_trip_and_canonical_layout(nb):
    for f in ("s2i.nnspm", "vad.nnspm", "kws_galaxy.nnspm", "vad_acc32.nnspm"):
        raw = open(os.path.join(nb.MODEL_DIR, f), "rb").read()
        m = nb.Model.from_blob(raw)
        assert m.to_blob() == raw, f
    assert nb.Model.from_blob(os.path.join(nb.MODEL_DIR, "vad_acc32.nnspm")).acc32 is True
    m = nb.Model.from_blob(os.path.join(nb.MODEL_DIR, "vad.nnspm"))
    assert (m.nn_id, m.numlayers, m.size_layer, m.acc32) == (1, 5, [240, 28, 28, 28, 28, 2], False)
    m.set_acc32(True)
    assert m.to_blob() == open(os.path.join(nb.MODEL_DIR, "vad_acc32.nnspm"), "rb").read()


def test_blob_rejects_garbage(nb):
    raw = bytearray(open(os.path.join(nb.MODEL_DIR, "vad.nnspm"), "rb").read())
    for bad in (b"", bytes(raw[:100]), b"XXXXXXXX" + bytes(raw[8:]), bytes(raw[:-64])):
        with pytest.raises(nb.NnspError):
            nb.Model.from_blob(bad)
    raw2 = bytearray(raw)
    raw2[12:16] = (77).to_bytes(4, "little")            # numlayers = 77
    with pytest.raises(nb.NnspError):
        nb.Model.from_blob(bytes(raw2))


def _py_interleave(mat):
    """python/nnsp_pack/c_weight_man.py:5-47 (arm_M4=True), restated for the test."""
    out = []
    rows, cols = mat.shape

    def block(sub):
        k, n = sub.shape
        for cp in range(n >> 1):
            for rp in range(k >> 1):
                out.extend(sub[2 * rp:2 * rp + 2, 2 * cp:2 * cp + 2].T.flatten())
            if k & 1:
                out.extend(sub[-1, 2 * cp:2 * cp + 2])
        if n & 1:
            out.extend(sub[:, -1])
    r = 0
    while r + 4 <= rows:
        block(mat[r:r + 4]); r += 4
    if rows - r:
        block(mat[r:])
    return np.array(out, np.int8)


@pytest.mark.parametrize("rows,cols", [(4, 8), (5, 8), (6, 7), (7, 9), (1, 5), (2, 2), (3, 1), (41, 72), (9, 241)])
def test_arm_interleave_remainder_rows_and_odd_columns(nb, rows, cols):
    """1/2/3-row remainder groups and the odd last column (affine.c:103-184), which no shipped model exercises."""
    lib = C.CDLL(os.path.join(ROOT, "nnsp_b200", "libnnsp_b200.so"))
    lib.nnsp_interleave_arm.restype = C.c_size_t
    lib.nnsp_deinterleave_arm.restype = C.c_size_t
    lib.nnsp_interleave_arm.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.nnsp_deinterleave_arm.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    rng = np.random.default_rng(rows * 100 + cols)
    mat = rng.integers(-128, 128, (rows, cols)).astype(np.int8)
    tab = np.zeros(rows * cols, np.int8)
    n = lib.nnsp_interleave_arm(mat.ctypes.data_as(C.c_void_p), rows, cols, tab.ctypes.data_as(C.c_void_p))
    assert n == rows * cols and (tab == _py_interleave(mat)).all()
    back = np.zeros_like(mat)
    assert lib.nnsp_deinterleave_arm(tab.ctypes.data_as(C.c_void_p), rows, cols, back.ctypes.data_as(C.c_void_p)) == rows * cols
    assert (back == mat).all()


def test_synth_is_deterministic(nb):
    a, b = nb.synth_pcm(20, 30, first_stream=3), nb.synth_pcm(20, 30, first_stream=3)
    assert (a == b).all() and a.dtype == np.int16
    assert (nb.synth_pcm(4, 30, first_stream=19) == a[16:20]).all()     # stream identity does not depend on the batch
    assert (a[14] == 0).all()                                           # stream 17: digital silence
