#!/bin/bash
# GPU call: warp-per-theta pixel kernel: parity + bench; LOS / chord occupancy variants.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_balmer_$name.log 2>&1
  env "$@" timeout 300 python bench.py --workload multi_species --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_multi_$name.log 2>&1
  python - <<PY
import json
for f in ["gpurun_out/bench_balmer_$name.log","gpurun_out/bench_multi_$name.log"]:
    for line in open(f):
        if line.startswith("{"):
            d=json.loads(line); r=d["roofline"]
            print("$name", f.split("_")[2], round(d["value"]), d["gpu_launches"], {k:round(v,4) for k,v in r["stage_ms"].items()})
PY
}
run base BAYSAR_LOS_MINB=4
run los5 BAYSAR_LOS_MINB=5
run los6 BAYSAR_LOS_MINB=6
run chord4 BAYSAR_CHORD_MINB=4
