#!/bin/bash
# GPU call: parity suite + both bench lines (no CPU baseline) after a LOS-stage change.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_balmer.log 2>&1
timeout 600 python bench.py --workload multi_species --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_multi.log 2>&1
python - <<PY
import json
for f in ["gpurun_out/bench_balmer.log","gpurun_out/bench_multi.log"]:
    for line in open(f):
        if line.startswith("{"):
            d=json.loads(line); r=d["roofline"]
            print(f, round(d["value"]), round(d["e2e"]["value"]), {k:round(v,4) for k,v in r["stage_ms"].items()})
PY
