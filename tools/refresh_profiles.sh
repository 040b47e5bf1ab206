#!/bin/bash
# Run ON THE GPU BOX (under gpurun): every measurement the round's profiles/ are built from, without a profiler.
#   bash tools/refresh_profiles.sh r2        -> gpurun_out/r2_*.{json,jsonl,txt}
# The ncu passes are separate gpurun calls (one profiler run per call): see DESIGN.md section 6 for the command lines.
tag=${1:-r2}
o=gpurun_out
python bench.py --impl reference --steps 20 --warmup 5 > $o/${tag}_bench_reference_arm.json 2> $o/${tag}_bench_reference_arm.err
python bench.py --steps 20 --warmup 5 > $o/${tag}_bench.json 2> $o/${tag}_bench.err
echo "bench rc=$?"
python tools/bench_configs.py cascade:8192 vad:4096 kws:16384:acc32 s2i:32768 --paths=split > $o/${tag}_configs.jsonl 2>&1
python tools/casc_timeline.py > $o/${tag}_cascade_timeline.txt 2>&1
for n in 72 64 28; do TC5_PROF=1 tools/tc5_gemm_bench 32768 $n; done > $o/${tag}_tc5_gemm_bench.txt 2>&1
tools/umma_rate > $o/${tag}_umma_rate.txt 2>&1
python tools/h2d_bw.py 262 > $o/${tag}_h2d_bw.txt 2>&1
