#!/bin/bash
# Turn the files a measurement run left in gpurun_out/ (see DESIGN.md section 6 for the commands) into the committed
# summaries under profiles/. usage: bash tools/refresh_profiles.sh v4
set -e
tag=${1:-v4}
cp gpurun_out/${tag}_bench.json profiles/r1_${tag}_bench.json
cp gpurun_out/${tag}_ref.json profiles/r1_${tag}_bench_reference_arm.json
cp gpurun_out/${tag}_configs.jsonl profiles/r1_${tag}_configs.jsonl
cp gpurun_out/${tag}_launches.csv profiles/r1_${tag}_launches.csv
{ echo "# ncu --metrics gpu__time_duration.sum --clock-control none over \`python bench.py --steps 2 --warmup 3\` (profiles/r1_${tag}_launches.csv)"
  python tools/step_share.py gpurun_out/${tag}_launches.csv
  python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
k = d["kernel_ms"]
print("bench.py (same code, no profiler): %.0f us per step with the two-stream pipeline; feat_kernel %.0f us = %.0f%% of (feat %.0f + network %.0f) from the CUDA events inside bench.py" % (
    1e3 * d["ms_per_step"], 1e3 * k["feat_kernel"], 100 * k["feat_kernel"] / (k["feat_kernel"] + k["network_kernels"]), 1e3 * k["feat_kernel"], 1e3 * k["network_kernels"]))
PY
} > profiles/r1_${tag}_step_share.txt
if [ -f gpurun_out/${tag}_full.ncu-rep ]; then
  { python tools/ncu_summary.py gpurun_out/${tag}_full.ncu-rep
    for k in feat_kernel "seg_kernel<(int)1>" scan_kernel "seg_kernel<(int)0>"; do python tools/ncu_hot.py gpurun_out/${tag}_full.ncu-rep "$k" 12; done
    for k in feat_kernel "seg_kernel<(int)1>" scan_kernel "seg_kernel<(int)0>"; do python tools/ncu_blocks.py gpurun_out/${tag}_full.ncu-rep "$k" 100 | head -2; done
  } > profiles/r1_${tag}_ncu_full.txt 2>&1
fi
