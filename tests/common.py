"""Shared helpers for the parity tests."""
import hashlib
import os

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
