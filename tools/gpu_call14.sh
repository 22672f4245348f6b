#!/bin/bash
# GPU call: folded pixel contraction: unit test first (short timeout), trace at bench size, then parity suite + bench.
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_pix_contract.py -m gpu -x -q -s > gpurun_out/pytest_pxc.log 2>&1; rc=$?; echo "pytest pxc rc=$rc" >> gpurun_out/pytest_pxc.log
grep "TN=\|passed\|failed\|rc=" gpurun_out/pytest_pxc.log | cut -c1-250
if [ $rc -eq 124 ]; then echo "HUNG"; exit 1; fi
BAYSAR_PXC_TRACE=1 timeout 120 python tools/pxc_probe.py > gpurun_out/pxc_probe.log 2>&1; cat gpurun_out/pxc_probe.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
BAYSAR_PXC=64 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu_tn64.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_tn64.log
tail -5 gpurun_out/pytest_gpu_tn64.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_balmer.log 2>&1; echo "rc=$?" >> gpurun_out/bench_balmer.log
BAYSAR_PXC=64 timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_balmer_tn64.log 2>&1; echo "rc=$?" >> gpurun_out/bench_balmer_tn64.log
python - <<PY
import json
for f in ["gpurun_out/bench_balmer.log","gpurun_out/bench_balmer_tn64.log"]:
    for line in open(f):
        if line.startswith("{"):
            d=json.loads(line); r=d["roofline"]
            print(f, round(d["value"]), {k:round(v,4) for k,v in r["stage_ms"].items()})
PY
