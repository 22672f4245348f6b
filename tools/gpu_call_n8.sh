#!/bin/bash
# 8-GPU call: weak-scaling bench lines (torchrun, NCCL) at N = 4 and 8 and the config-4 ensemble on 8 GPUs.
mkdir -p gpurun_out
N=${1:-8}
for n in 4 $N; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n${n}_balmer.log 2>&1; echo "rc=$?" >> gpurun_out/bench_n${n}_balmer.log
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --steps 20 --warmup 3 --workload multi_species --no-cpu-baseline > gpurun_out/bench_n${n}_multi.log 2>&1; echo "rc=$?" >> gpurun_out/bench_n${n}_multi.log
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 tools/ensemble_bench.py --walkers-per-gpu 131072 --steps 2 > gpurun_out/ensemble_n${N}.log 2>&1; echo "rc=$?" >> gpurun_out/ensemble_n${N}.log
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_n[48]_*.log")) + sorted(glob.glob("gpurun_out/ensemble_n[48].log")):
    for line in open(f):
        if line.startswith("{"):
            d = json.loads(line)
            print(f, d["n_gpus"], round(d["value"]), round(d.get("e2e", {}).get("value", 0)), d.get("ms_per_step"))
PY
