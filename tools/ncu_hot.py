#!/usr/bin/env python
"""Hot source lines of a kernel from an .ncu-rep (--set full --import-source on; code built with -lineinfo).
Joins ncu's per-SASS-instruction samples with nvdisasm's line table of the in-tree library.
usage: python tools/ncu_hot.py rep.ncu-rep kernel-substring [top-n] [lib.so]"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

rep, want = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
lib = sys.argv[4] if len(sys.argv) > 4 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "nnsp_b200", "libnnsp_b200.so")

# 1. line table: function -> {offset: (file, line)}
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
linemap = {}
for cub in glob.glob(os.path.join(tmp, "*.cubin")):
    txt = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
    fn, cur = None, None
    for ln in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            fn = m.group(1); linemap[fn] = {}; cur = None; continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", ln)
        if m and fn:
            linemap[fn][int(m.group(1), 16)] = (cur, m.group(2).strip())
demangle = {}
for fn in linemap:
    demangle[fn] = subprocess.run(["cu++filt", fn], capture_output=True, text=True).stdout.strip()

# 2. samples per SASS instruction
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
i, seen = 0, set()
while i < len(lines):
    if lines[i].startswith('"Kernel Name"'):
        kname = next(csv.reader([lines[i]]))[1]
        j = i + 1
        block = []
        while j < len(lines) and not lines[j].startswith('"Kernel Name"'):
            block.append(lines[j]); j += 1
        i = j
        if want not in kname or kname in seen:
            continue
        seen.add(kname)
        rows = list(csv.reader(block))
        hdr = rows[0]
        ca, cs, ci = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
        stall_cols = [(k, h) for k, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        cx = hdr.index("L1 Wavefronts Shared Excessive") if "L1 Wavefronts Shared Excessive" in hdr else None
        cw = hdr.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in hdr else None
        exc_line, wav_line = collections.Counter(), collections.Counter()
        norm = lambda x: x.replace("(bool)1", "true").replace("(bool)0", "false").replace("void ", "").replace(" ", "")
        fn = [f for f, d in demangle.items() if norm(d) == norm(kname)]
        fn = fn[0] if fn else None
        base = int(rows[1][ca], 16)
        by_line, inst_line, stall_line = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
        tot, tot_inst = 0, 0
        for r in rows[1:]:
            off = int(r[ca], 16) - base
            s = int(float(r[cs] or 0)); ie = int(float(r[ci] or 0))
            loc = linemap.get(fn, {}).get(off, (None, ""))[0] if fn else None
            by_line[loc] += s; inst_line[loc] += ie; tot += s; tot_inst += ie
            if cx is not None:
                exc_line[loc] += int(float(r[cx] or 0)); wav_line[loc] += int(float(r[cw] or 0))
            for k, h in stall_cols:
                v = int(float(r[k] or 0))
                if v: stall_line[loc][h] += v
        print("== %s: %d samples, %d warp-instructions executed" % (kname, tot, tot_inst))
        for loc, v in by_line.most_common(topn):
            st = ", ".join("%s %d" % (h.replace("stall_", ""), c) for h, c in stall_line[loc].most_common(3))
            print("%6.2f%% samples %6.2f%% inst  %s:%s   [%s]" % (100.0 * v / max(tot, 1), 100.0 * inst_line[loc] / max(tot_inst, 1),
                                                              loc[0] if loc else "?", loc[1] if loc else "?", st))
        if cx is not None and sum(exc_line.values()):
            print("   shared-memory wavefronts: %d, of which excessive (bank conflicts) %d; worst lines:" % (sum(wav_line.values()), sum(exc_line.values())))
            for loc, v in exc_line.most_common(8):
                if v: print("     %10d excessive of %10d  %s:%s" % (v, wav_line[loc], loc[0] if loc else "?", loc[1] if loc else "?"))
    else:
        i += 1
