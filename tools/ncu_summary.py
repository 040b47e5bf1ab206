#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the handful of metrics DESIGN.md / profiles/ quote.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_kernel.txt"""
import csv
import subprocess
import sys

WANT = """gpu__time_duration.sum launch__grid_size launch__block_size launch__registers_per_thread
launch__occupancy_limit_registers launch__occupancy_limit_shared_mem sm__warps_active.avg.per_cycle_active
sm__warps_active.avg.pct_of_peak_sustained_active smsp__issue_active.avg.per_cycle_active
smsp__warps_eligible.avg.per_cycle_active smsp__inst_executed.sum smsp__thread_inst_executed_per_inst_executed.ratio
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
sm__throughput.avg.pct_of_peak_sustained_elapsed gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
dram__bytes_read.sum dram__bytes_write.sum lts__t_bytes.sum l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum sm__cycles_elapsed.max""".split()


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("kernel:", d.get("Kernel Name", "?"))
        for w in WANT:
            if w in d:
                print("  %-72s %s %s" % (w, d[w], units[hdr.index(w)]))
        stalls = sorted(((float(v.replace(",", "")), k) for k, v in d.items()
                         if "issue_stalled" in k and k.endswith("_per_warp_active.pct") and v), reverse=True)
        for v, k in stalls[:6]:
            print("  %-72s %.1f %%" % (k, v))


if __name__ == "__main__":
    main(sys.argv[1])
