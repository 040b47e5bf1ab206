#!/usr/bin/env python
"""Per-kernel share of the device-resident step from an ncu launch list
(ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv python bench.py --steps 2 --warmup 3).
usage: python tools/step_share.py launches.csv [launches-per-step=7] [steps=5]"""
import collections
import csv
import sys

path = sys.argv[1]
per_step = int(sys.argv[2]) if len(sys.argv) > 2 else 7
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
rows = list(csv.DictReader(l for l in open(path) if not l.startswith("==")))
launches = [(int(r["ID"]), r["Kernel Name"].split("(")[0].replace("void ", ""), float(r["Metric Value"].replace(",", "")))
            for r in rows if r["Metric Name"] == "gpu__time_duration.sum"]
# the device-resident steps are the first `steps` groups that start with feat_kernel after the one-off reset kernel
first = next(i for i, l in enumerate(launches) if l[1].startswith("feat_kernel"))
sel = launches[first:first + per_step * steps]
agg = collections.OrderedDict()
for _, name, ns in sel:
    agg.setdefault(name, []).append(ns)
tot = sum(sum(v) for v in agg.values()) / steps
print("# device-resident steps: launches %d..%d of the list (%d steps x %d launches); later launches are the host-buffer calls" %
      (sel[0][0], sel[-1][0], steps, per_step))
for name, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("%-28s launches=%2d avg=%7.1f us share of the serialised step=%5.1f%%" % (name, len(v), sum(v) / len(v) / 1e3, 100.0 * sum(v) / steps / tot))
print("sum per step = %.1f us (serialised under ncu, cold cache)" % (tot / 1e3))
