#!/usr/bin/env python
"""Dynamic instruction mix of a kernel from an .ncu-rep (--set full --import-source on): opcode histogram and the
basic blocks (runs of SASS instructions with the same execution count) that carry the most executed instructions.
usage: python tools/ncu_blocks.py rep.ncu-rep kernel-substring [min-percent]"""
import collections
import csv
import subprocess
import sys

rep, want = sys.argv[1], sys.argv[2]
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
lines = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
i, seen = 0, set()
while i < len(lines):
    if not lines[i].startswith('"Kernel Name"'):
        i += 1
        continue
    kname = next(csv.reader([lines[i]]))[1]
    j, block = i + 1, []
    while j < len(lines) and not lines[j].startswith('"Kernel Name"'):
        block.append(lines[j]); j += 1
    i = j
    if want not in kname or kname in seen:
        continue
    seen.add(kname)
    rows = list(csv.reader(block))
    hdr = rows[0]
    ci, cs = hdr.index("Instructions Executed"), hdr.index("Source")
    ins = []
    for r in rows[1:]:
        try:
            n = int(r[ci])
        except (ValueError, IndexError):
            continue
        t = r[cs].strip().split()
        op = (t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?"))
        ins.append((n, op))
    tot = sum(n for n, _ in ins)
    print("== %s: %d warp-instructions" % (kname, tot))
    mix = collections.Counter()
    for n, op in ins:
        mix[op.split(".")[0]] += n
    print("   mix: " + ", ".join("%s %.1f%%" % (k, 100.0 * v / tot) for k, v in mix.most_common(18)))
    prev, acc = None, []
    def flush():
        if acc and 100.0 * prev * len(acc) / tot >= minpct:
            print("   %5.1f%%  x%-9d %3d instr: %s" % (100.0 * prev * len(acc) / tot, prev, len(acc), " ".join(acc)))
    for n, op in ins:
        if n != prev:
            flush()
            prev, acc = n, []
        acc.append(op)
    flush()
