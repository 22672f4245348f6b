#!/bin/bash
# GPU call: parity + bench with chord register-budget variants (multi-species) and the LOS changes.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for mb in 3 4 5; do
  BAYSAR_CHORD_MINB=$mb timeout 300 python bench.py --workload multi_species --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_multi_cminb$mb.log 2>&1
  python - <<PY
import json
for f in ["gpurun_out/bench_multi_cminb$mb.log"]:
    for line in open(f):
        if line.startswith("{"):
            d=json.loads(line); r=d["roofline"]
            print("chord minb=$mb", round(d["value"]), {k:round(v,4) for k,v in r["stage_ms"].items()})
PY
done
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_balmer.log 2>&1
python - <<PY
import json
for line in open("gpurun_out/bench_balmer.log"):
    if line.startswith("{"):
        d=json.loads(line); r=d["roofline"]
        print("balmer", round(d["value"]), {k:round(v,4) for k,v in r["stage_ms"].items()})
PY
