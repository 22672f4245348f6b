#!/bin/bash
# 2-GPU call: overlapped vs synchronous logP all-gather in the bench loop.
mkdir -p gpurun_out
for mode in 0 1; do
  BAYSAR_BENCH_SYNC_GATHER=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2957$mode bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n2_balmer_sync$mode.log 2>&1; echo "rc=$?" >> gpurun_out/bench_n2_balmer_sync$mode.log
done
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1_balmer.log 2>&1
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_n2_balmer_sync*.log")) + ["gpurun_out/bench_n1_balmer.log"]:
    for line in open(f):
        if line.startswith("{"):
            d = json.loads(line)
            print(f, d["n_gpus"], round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], d["config"].get("logp_all_gather"))
    print(open(f).read()[-300:] if "rc=0" not in open(f).read() and "n1" not in f else "")
PY
