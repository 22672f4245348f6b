#!/bin/bash
# GPU call: parity + bench with LOS register-cap variants.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for mb in 4 5 6; do
  BAYSAR_LOS_MINB=$mb timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_balmer_minb$mb.log 2>&1
  BAYSAR_LOS_MINB=$mb timeout 300 python bench.py --workload multi_species --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_multi_minb$mb.log 2>&1
  python - <<PY
import json
for f in ["gpurun_out/bench_balmer_minb$mb.log","gpurun_out/bench_multi_minb$mb.log"]:
    for line in open(f):
        if line.startswith("{"):
            d=json.loads(line); r=d["roofline"]
            print("minb=$mb", f, round(d["value"]), {k:round(v,4) for k,v in r["stage_ms"].items()})
PY
done
