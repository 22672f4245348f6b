#!/bin/bash
# GPU call: parity tests, bench lines (balmer_series, multi_species), ncu launch list and full captures.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/gpu.txt
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_balmer.log 2>&1; echo "rc=$?" >> gpurun_out/bench_balmer.log
python bench.py --workload multi_species --steps 20 --warmup 3 > gpurun_out/bench_multi.log 2>&1; echo "rc=$?" >> gpurun_out/bench_multi.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:chord_stage_kernel|if_contract_kernel|los_stage_kernel|pixel_stage_kernel' -s 12 -c 4 -f -o gpurun_out/prof_r1s2 $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/bench_balmer.log | cut -c1-600; tail -2 gpurun_out/bench_multi.log | cut -c1-600; tail -3 gpurun_out/ncu_full.log
