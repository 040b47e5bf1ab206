#!/usr/bin/env python
"""End-to-end rate of the batched NNSPClass through the host-buffer calls (pinned host PCM, H2D + kernels + D2H inside the
timed region, two buffer pairs as in INTEGRATION.md). usage: python tools/e2e_batch.py [model=vad] [streams=4096] [frames=100] [calls=40]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import nnsp_b200 as nb  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "vad"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
T = int(sys.argv[3]) if len(sys.argv) > 3 else 100
N = int(sys.argv[4]) if len(sys.argv) > 4 else 40
FILES = {"s2i": "s2i.nnspm", "vad": "vad.nnspm", "kws": "kws_galaxy.nnspm"}
b = nb.NNSPBatch(nb.Model.from_blob(os.path.join(nb.MODEL_DIR, FILES[name])), S)
pool = nb.synth_pcm(min(S, 1024), T)
pin = [nb.PinnedArray((S, T * 160), np.int16) for _ in range(2)]
res = [nb.PinnedArray((S, T), nb.RESULT_DT) for _ in range(2)]
for p in pin:
    p.array[...] = np.tile(pool, ((S + len(pool) - 1) // len(pool), 1))[:S]


def loop(n):
    prev = None
    for k in range(n):
        tk = b.exec_host_async(pin[k & 1].array, res[k & 1].array)
        if prev is not None:
            b.wait_host(prev)
        prev = tk
    b.wait_host(prev)


loop(6)
t0 = time.perf_counter()
loop(N)
dt = (time.perf_counter() - t0) / N
print("%s x %d streams x %d frames per call: %.3f ms per call end to end, %.2f M audio-s/s, H2D %.1f GB/s" % (
    name, S, T, dt * 1e3, S * T * 0.01 / dt / 1e6, S * T * 320 / dt / 1e9))
