#!/bin/bash
# GPU call: LOS stage at 7 / 8 CTAs per SM (single wave at 4096 thetas) on the Balmer demo.
mkdir -p gpurun_out
for mb in 4 7 8; do
  BAYSAR_LOS_MINB=$mb timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_balmer_los$mb.log 2>&1
  python - <<PY
import json
for line in open("gpurun_out/bench_balmer_los$mb.log"):
    if line.startswith("{"):
        d=json.loads(line); r=d["roofline"]
        print("los minb=$mb", round(d["value"]), {k:round(v,4) for k,v in r["stage_ms"].items()})
PY
done
