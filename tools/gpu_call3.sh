#!/bin/bash
# GPU call: parity tests + bench lines for both workloads (no ncu).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_balmer.log 2>&1; echo "rc=$?" >> gpurun_out/bench_balmer.log
python bench.py --workload multi_species --steps 20 --warmup 3 > gpurun_out/bench_multi.log 2>&1; echo "rc=$?" >> gpurun_out/bench_multi.log
tail -4 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/bench_balmer.log | cut -c1-300; tail -2 gpurun_out/bench_multi.log | cut -c1-300
