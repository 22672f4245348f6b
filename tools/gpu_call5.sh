#!/bin/bash
# GPU call: new TMEM-fed contraction kernel: unit test first (short timeout), then parity suite and bench lines.
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_if_contract.py -m gpu -x -q -s > gpurun_out/pytest_ifc.log 2>&1; rc=$?; echo "pytest ifc rc=$rc" >> gpurun_out/pytest_ifc.log
tail -15 gpurun_out/pytest_ifc.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 120 python tools/ifc_probe.py 4096 8002 > gpurun_out/ifc_probe.log 2>&1; tail -3 gpurun_out/ifc_probe.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_balmer.log 2>&1; echo "rc=$?" >> gpurun_out/bench_balmer.log
timeout 600 python bench.py --workload multi_species --steps 20 --warmup 3 > gpurun_out/bench_multi.log 2>&1; echo "rc=$?" >> gpurun_out/bench_multi.log
tail -4 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/bench_balmer.log | cut -c1-300; tail -2 gpurun_out/bench_multi.log | cut -c1-300
