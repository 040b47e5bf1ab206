#!/usr/bin/env python
"""One device-resident step of each configuration from an ncu launch list of tools/bench_configs.py (metrics:
gpu__time_duration.sum, smsp__inst_executed.sum, smsp__issue_active..., sm__warps_active..., sm__pipe_tensor_cycles_active...,
dram__bytes_read/write.sum). usage: python tools/launch_table.py launches.csv "title 1" "title 2" ..."""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
data = collections.OrderedDict()
for row in csv.DictReader(lines):
    key = (int(row["ID"]), row["Kernel Name"].split("(")[0].replace("void ", ""), row["Grid Size"], row["Block Size"])
    data.setdefault(key, {})[row["Metric Name"]] = row["Metric Value"]
items = list(data.items())
titles = sys.argv[2:]
n = len(items)
per = n // len(titles)
print("# one device-resident step (7 launches) of each configuration; times are serialised, cold-cache")
for t, title in enumerate(titles):
    hi = per * (t + 1) - 1                      # the last launch of a configuration is the lone call that reads the kernel times
    lo = hi - 7
    print(title)
    tot = sum(float(v["gpu__time_duration.sum"]) for _, v in items[lo:hi])
    for k, v in items[lo:hi]:
        us = float(v["gpu__time_duration.sum"])
        print("  %-22s grid %-14s block %-12s %9.1f us %5.1f%%  warp-inst %12s  issue/clk/SMSP %s  tensor-pipe %5s%%  warps active %5s%%  dram rd %6.1f MB wr %6.1f MB" % (
            k[1][:22], k[2], k[3], us / 1e3, 100 * us / tot, v["smsp__inst_executed.sum"], v["smsp__issue_active.avg.per_cycle_active"],
            v["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"], v["sm__warps_active.avg.pct_of_peak_sustained_active"],
            float(v["dram__bytes_read.sum"]) / 1e6, float(v["dram__bytes_write.sum"]) / 1e6))
    print("  sum %.1f us" % (tot / 1e3))
