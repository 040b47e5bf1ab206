#!/usr/bin/env python
"""One steady-state cascade call out of an ncu launch list (gpu__time_duration.sum + a few counters per launch, CSV):
the launches between two feat_kernel launches, serialised and cold-cache -- compare SHARES, not absolutes.
usage: python tools/step_launches.py launches.csv"""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
data = collections.OrderedDict()
for row in csv.DictReader(lines):
    key = (int(row["ID"]), row["Kernel Name"].split("(")[0].replace("void ", "").replace("nnsp::", ""), row["Grid Size"], row["Block Size"])
    data.setdefault(key, {})[row["Metric Name"]] = row["Metric Value"]
items = list(data.items())
feat = [i for i, (k, _) in enumerate(items) if k[1].startswith("feat_kernel")]
lo, hi = feat[-2], feat[-1]
tot = sum(float(v["gpu__time_duration.sum"]) for _, v in items[lo:hi]) / 1e3
print("# %d launches of one call; serialised sum %.1f us" % (hi - lo, tot))
for k, v in items[lo:hi]:
    us = float(v["gpu__time_duration.sum"]) / 1e3
    print("  %-30s grid %-13s block %-12s %8.1f us %5.1f%%  warp-inst %11s  issue/clk/SMSP %4s  tensor %5s%%  warps active %5s%%  dram rd %6.1f wr %6.1f MB" % (
        k[1][:30], k[2], k[3], us, 100 * us / tot, v.get("smsp__inst_executed.sum", "-"), v.get("smsp__issue_active.avg.per_cycle_active", "-"),
        v.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "-"), v.get("sm__warps_active.avg.pct_of_peak_sustained_active", "-"),
        float(v.get("dram__bytes_read.sum", 0)) / 1e6, float(v.get("dram__bytes_write.sum", 0)) / 1e6))
