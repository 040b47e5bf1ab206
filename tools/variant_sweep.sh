#!/bin/bash
# Measurement helper (GPU box): tools/bench_configs.py under kernel variants (tools/build_variant.sh) and run-time knobs.
out=${1:-gpurun_out/sweep.jsonl}
: > $out
CFG="cascade:8192 vad:4096 kws:16384:acc32 s2i:32768"
run() { echo "# $1" >> $out; env $2 python tools/bench_configs.py $CFG --paths=split --cascade-paths=sorted >> $out 2>&1; }
run product ""
for v in build/variants/*.so; do run "$v" "NNSP_B200_LIB=$PWD/$v"; done
