#!/bin/bash
# GPU call: LOS team mode (one CTA per theta for long lines of sight): parity, benches, shape sweep with two thresholds.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
BAYSAR_LOS_TEAM_L=40 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden_fixtures" > gpurun_out/pytest_team.log 2>&1; tail -2 gpurun_out/pytest_team.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_balmer.log 2>&1
timeout 600 python bench.py --workload multi_species --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_multi.log 2>&1
BAYSAR_LOS_TEAM_L=90 timeout 600 python bench.py --workload multi_species --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_multi_team.log 2>&1
python - <<PY
import json
for f in ["gpurun_out/bench_balmer.log","gpurun_out/bench_multi.log","gpurun_out/bench_multi_team.log"]:
    for line in open(f):
        if line.startswith("{"):
            d=json.loads(line); r=d["roofline"]
            print(f, round(d["value"]), {k:round(v,4) for k,v in r["stage_ms"].items()})
PY
timeout 900 python tools/shape_sweep.py --out gpurun_out/shape_sweep.jsonl > gpurun_out/shape_sweep.log 2>&1
BAYSAR_LOS_TEAM_L=150 timeout 900 python tools/shape_sweep.py --out gpurun_out/shape_sweep_team150.jsonl > gpurun_out/shape_sweep_team150.log 2>&1
python - <<PY
import json
for f in ["gpurun_out/shape_sweep.jsonl","gpurun_out/shape_sweep_team150.jsonl"]:
    print(f)
    for line in open(f):
        d=json.loads(line); print(" ", d["L"], d["n_G"], d["P"], d["n_B"], round(d["evals_per_s"]), round(d["stage_ms"]["los"],2), round(d["stage_ms"]["chord"],2))
PY
