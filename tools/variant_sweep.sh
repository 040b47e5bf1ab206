#!/bin/bash
# Measurement helper (GPU box): the configurations of tools/bench_configs.py under every kernel variant built by
# tools/build_variant.sh, plus run-time knobs. Output: gpurun_out/$1
out=${1:-gpurun_out/sweep.jsonl}
: > $out
CFG="cascade:8192 vad:4096 kws:16384:acc32 s2i:32768"
run() { echo "# $1" >> $out; env $2 python tools/bench_configs.py $CFG --paths=split --cascade-paths=sorted >> $out 2>&1; }
run product ""
for v in build/variants/*.so; do run "$v" "NNSP_B200_LIB=$PWD/$v"; done
run "product pad40k" "NNSP_B200_FEAT_SMEM_PAD=40000"
run "seg96 pad40k" "NNSP_B200_LIB=$PWD/build/variants/libnnsp_b200_seg96.so NNSP_B200_FEAT_SMEM_PAD=40000"
run "product rounds2" "NNSP_B200_CASCADE_ROUNDS=2"
