#!/bin/bash
# Measurement helper (GPU box): cascade timeline + tools/bench_configs.py under kernel variants (tools/build_variant.sh).
out=${1:-gpurun_out/sweep.txt}
: > $out
run() { echo "## $1" >> $out; env $2 python tools/casc_timeline.py 2>&1 | tail -1 >> $out; env $2 python tools/bench_configs.py cascade:8192 kws:16384:acc32 --paths=split --cascade-paths=sorted >> $out 2>&1; }
run product ""
for v in build/variants/*.so; do run "$v" "NNSP_B200_LIB=$PWD/$v"; done
