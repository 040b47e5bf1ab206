"""Shared helpers for the parity tests."""
import hashlib
import os

import struct

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")
MODEL_NAME = {0: "s2i", 1: "vad", 2: "kws"}
TAP_NAMES = ["logmel", "feat", "act", "logits", "h", "c", "post"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def golden():
    return np.load(GOLDEN)


def res4(res):
    """structured result array -> int16 [T, 4] (trigger, outputs[3])"""
    return res.view(np.int16).reshape(len(res), 4)


def have_reference_tree():
    return os.path.isdir("/root/reference/ns-nnsp/src")


def make_blob(nn_id, sizes, types, acts, qk, qi, qb, seed=0):
    """NNSPM1 container with random table-layout weights: any byte string is a valid table."""
    rng = np.random.default_rng(seed)
    nl = len(types)
    hdr = b"NNSPM1\0\0" + struct.pack("<ii", nn_id, nl)
    sl = list(sizes) + [0] * (11 - len(sizes))
    hdr += struct.pack("<11h", *sl) + b"\0\0"
    hdr += rng.integers(-200000, 200000, 40, dtype=np.int32).tobytes() + rng.integers(1, 40000, 40, dtype=np.int32).tobytes()
    recs, body = b"", b""
    for i in range(10):
        if i < nl:
            rows, cols = sizes[i + 1], sizes[i]
            nr = 4 * rows if types[i] == 1 else rows
            kb, rb, bc = nr * cols, (nr * rows if types[i] == 1 else 0), nr
            qin = qi[i + 1] if i + 1 < nl else 0
            recs += struct.pack("<10i", types[i], acts[i], qk[i], qi[i], qb[i], 0, kb, rb, bc, qin)
            for n in (kb, rb):
                a = rng.integers(-128, 128, n, dtype=np.int8).tobytes()
                body += a + b"\0" * (-len(a) % 4)
            a = rng.integers(-32768, 32768, bc, dtype=np.int16).tobytes()
            body += a + b"\0" * (-len(a) % 4)
        else:
            recs += b"\0" * 40
    return hdr + recs + body
