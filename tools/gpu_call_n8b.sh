#!/bin/bash
# 8-GPU call: bench lines with the overlapped logP all-gather.
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 8 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n8_balmer.log 2>&1; echo "rc=$?" >> gpurun_out/bench_n8_balmer.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29582 bench.py --gpus 8 --steps 20 --warmup 3 --workload multi_species --no-cpu-baseline > gpurun_out/bench_n8_multi.log 2>&1; echo "rc=$?" >> gpurun_out/bench_n8_multi.log
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_n8_*.log")):
    for line in open(f):
        if line.startswith("{"):
            d = json.loads(line)
            print(f, d["n_gpus"], round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], d["config"].get("logp_all_gather"))
PY
