#!/bin/bash
# GPU call: quick loop for the folded contraction: unit test, traced probe, parity file, bench line.
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_pix_contract.py -m gpu -x -q > gpurun_out/pytest_pxc.log 2>&1; rc=$?; tail -2 gpurun_out/pytest_pxc.log
if [ $rc -ne 0 ]; then tail -30 gpurun_out/pytest_pxc.log; exit 1; fi
BAYSAR_PXC_TRACE=1 timeout 120 python tools/pxc_probe.py > gpurun_out/pxc_probe.log 2>&1; grep -v "^TN=64\|conv.total\|build.total\|epi.wait" gpurun_out/pxc_probe.log | cut -c1-150
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_balmer.log 2>&1; echo "rc=$?" >> gpurun_out/bench_balmer.log
python - <<PY
import json
for f in ["gpurun_out/bench_balmer.log"]:
    for line in open(f):
        if line.startswith("{"):
            d=json.loads(line); r=d["roofline"]
            print(f, round(d["value"]), {k:round(v,4) for k,v in r["stage_ms"].items()}, r["kernel"], round(r["frac"],3))
PY
