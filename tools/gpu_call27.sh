#!/bin/bash
# GPU call: tile-schedule chunking of the folded contraction (L2 residency of the spectra).
mkdir -p gpurun_out
for ck in 0 16 8 4; do
  BAYSAR_PXC_CHUNK=$ck timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_balmer_ck$ck.log 2>&1
  python - <<PY
import json
for line in open("gpurun_out/bench_balmer_ck$ck.log"):
    if line.startswith("{"):
        d=json.loads(line); r=d["roofline"]
        print("chunk=$ck", round(d["value"]), {k:round(v,4) for k,v in r["stage_ms"].items()})
PY
done
