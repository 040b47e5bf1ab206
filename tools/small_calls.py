#!/usr/bin/env python
"""Real-time regime: many short calls (T frames per call) on device-resident PCM. Prints device time per call (CUDA events
around the whole run) and host time per call (how fast the calls can be issued).
usage: python tools/small_calls.py [streams=4096] [frames_per_call=2] [calls=2000] [model=vad]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import nnsp_b200 as nb  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 2
N = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
name = sys.argv[4] if len(sys.argv) > 4 else "vad"
FILES = {"s2i": "s2i.nnspm", "vad": "vad.nnspm", "kws": "kws_galaxy.nnspm"}
m = nb.Model.from_blob(os.path.join(nb.MODEL_DIR, FILES[name]), acc32=False)
b = nb.NNSPBatch(m, S)
ring = 8                                                   # a ring of device PCM buffers, as an ingest stage would fill
pcm = nb.synth_pcm(S, T * ring)
d = [nb.DeviceArray.from_host(np.ascontiguousarray(pcm[:, k * T * 160:(k + 1) * T * 160])) for k in range(ring)]
res = [nb.DeviceArray((S, T), nb.RESULT_DT) for _ in range(ring)]
for k in range(64):
    b.exec_device(d[k % ring], T * 160, T, res[k % ring])
b.sync()
ev0, ev1 = nb.Event(), nb.Event()
ev0.record(b.stream)
t0 = time.perf_counter()
for k in range(N):
    b.exec_device(d[k % ring], T * 160, T, res[k % ring])
t_issue = time.perf_counter() - t0
b.sync()
t_all = time.perf_counter() - t0
ev1.record(b.stream)
b.sync()
print("%s x %d streams, %d frames per call, %d calls: %.1f us per call wall (%.1f us to issue), %.2f M audio-s/s, %.1f x real time per stream" % (
    name, S, T, N, 1e6 * t_all / N, 1e6 * t_issue / N, S * T * 0.01 * N / t_all / 1e6, T * 0.01 * N / t_all))
