#!/bin/bash
# GPU call: shape sweep after the eight-warp team for L > 800; parity of a forced-team run.
mkdir -p gpurun_out
timeout 900 python tools/shape_sweep.py --out gpurun_out/shape_sweep.jsonl > gpurun_out/shape_sweep.log 2>&1
python - <<PY
import json
for line in open("gpurun_out/shape_sweep.jsonl"):
    d=json.loads(line); print(" ", d["L"], d["n_G"], d["P"], d["n_B"], round(d["evals_per_s"]), round(d["stage_ms"]["los"],2), round(d["stage_ms"]["chord"],2), d["thetas_per_step"])
PY
