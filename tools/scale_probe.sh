#!/bin/bash
# Host topology of a multi-GPU box and the e2e leg of bench.py with and without NUMA-local host binding.
# usage (on the GPU box): bash tools/scale_probe.sh "8 4 2" ["0 1"] > gpurun_out/scale_probe.log
mkdir -p gpurun_out
nvidia-smi topo -m
lscpu | grep -i -E "numa|socket|^CPU\(s\)|model name"
echo "nproc $(nproc)  cpuset $(cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null)"
python - <<'PY'
import sys; sys.path.insert(0, ".")
from nnsp_b200.shard import local_cpus
import nnsp_b200 as nb
for d in range(nb.device_count()):
    c = sorted(local_cpus(d))
    print("gpu", d, "local cpus", len(c), (c[0], c[-1]) if c else None)
PY
port=29611
for n in $1; do
  for bind in ${2:-0 1}; do
    port=$((port+1))
    NNSP_BENCH_BIND=$bind python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
        bench.py --gpus $n --steps 20 --warmup 3 2>gpurun_out/scale_err_${n}_${bind}.log | tail -1 > gpurun_out/scale_${n}_${bind}.json
    python - <<PY
import json
d = json.loads(open("gpurun_out/scale_${n}_${bind}.json").read())
print("N=$n bind=$bind value %.2f M  e2e %.2f M (%.3f ms/step)  host: %s" % (d["value"]/1e6, d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"], d["config"]["host"]))
PY
  done
done
