"""Pinned host -> device copy rate of the link the e2e number is bound by: python tools/h2d_bw.py [MB]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nnsp_b200 as nb
from nnsp_b200.capi import lib, check
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 131
n = mb * 1000 * 1000
pin = nb.PinnedArray((n,), np.uint8)
pin.array[:] = 1
dev = nb.DeviceArray((n,), np.uint8)
L = lib()
for _ in range(3):
    check(L.nnsp_b200_memcpy_h2d(0, dev.ptr, pin.ptr, n))
best = 1e9
for _ in range(10):
    t0 = time.perf_counter(); check(L.nnsp_b200_memcpy_h2d(0, dev.ptr, pin.ptr, n)); best = min(best, time.perf_counter() - t0)
print("H2D %d MB pinned: %.3f ms = %.1f GB/s" % (mb, best * 1e3, n / best / 1e9))
out = np.empty(n // 40, np.uint8)
pin2 = nb.PinnedArray((n // 40,), np.uint8)
best = 1e9
for _ in range(10):
    t0 = time.perf_counter(); check(L.nnsp_b200_memcpy_d2h(0, pin2.ptr, dev.ptr, n // 40)); best = min(best, time.perf_counter() - t0)
print("D2H %.1f MB pinned: %.3f ms = %.1f GB/s" % (n / 40e6, best * 1e3, n / 40 / best / 1e9))
