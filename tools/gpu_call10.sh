#!/bin/bash
# GPU call: quick bench of both workloads (stage times).
mkdir -p gpurun_out
for wl in balmer_series multi_species; do
  timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick_$wl.log 2>&1
  python - <<PY
import json
for line in open("gpurun_out/bench_quick_$wl.log"):
    if line.startswith("{"):
        d=json.loads(line); r=d["roofline"]
        print("$wl", round(d["value"]), {k:round(v,4) for k,v in r["stage_ms"].items()}, "ok", d["status_ok_fraction"])
PY
done
