#!/usr/bin/env python
"""Device timeline of back-to-back cascade calls (nnsp_b200_cascade_timeline): when the front end and the controller /
network chain of consecutive calls ran, i.e. how far they overlap. usage: python tools/casc_timeline.py [streams] [frames]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import nnsp_b200 as nb  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
pool = min(S, 2048)
base = nb.synth_pcm(pool, T)
pcm = np.tile(base, ((S + pool - 1) // pool, 1))[:S]
d = [nb.DeviceArray.from_host(pcm), nb.DeviceArray.from_host(np.roll(pcm, 7, axis=0))]
models = [nb.Model.from_blob(os.path.join(nb.MODEL_DIR, f)) for f in ("s2i.nnspm", "vad.nnspm", "kws_galaxy.nnspm")]
h = nb.Cascade(models, S)
res = nb.DeviceArray((S, T), nb.CASCADE_RESULT_DT)
for i in range(30):
    h.exec_device(d[i & 1], T * 160, T, res)
h.sync()
for i in range(8):
    h.exec_device(d[i & 1], T * 160, T, res)
tl = h.timeline()
print("# %d streams x %d frames per call, lib %s; ms after the first call's front end started" % (S, T, os.path.basename(os.environ.get("NNSP_B200_LIB", "product"))))
print("# call   front-end start..done      chain start..done")
for k, r in enumerate(tl):
    print("  %d      %7.3f .. %7.3f        %7.3f .. %7.3f" % (k, r[0], r[1], r[2], r[3]))
print("# per call: %.3f ms" % ((tl[-1][3] - tl[0][3]) / (len(tl) - 1)))
