#!/bin/bash
# GPU call: back-off sleeps in the long mbarrier waits of the folded contraction.
mkdir -p gpurun_out
for cfg in "0 0" "200 0" "1000 0" "200 100" "1000 200"; do
  set -- $cfg
  BAYSAR_PXC_SLEEP_EPI=$1 BAYSAR_PXC_SLEEP_BUILD=$2 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_balmer_sl.log 2>&1
  python - <<PY
import json
for line in open("gpurun_out/bench_balmer_sl.log"):
    if line.startswith("{"):
        d=json.loads(line); r=d["roofline"]
        print("sleep epi=$1 build=$2", round(d["value"]), {k:round(v,4) for k,v in r["stage_ms"].items()})
PY
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "contraction_variants or tensor_core" > gpurun_out/pytest_variants.log 2>&1; tail -3 gpurun_out/pytest_variants.log
