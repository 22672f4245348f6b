#!/bin/bash
# GPU call: full evidence run: parity suite, bench lines (both workloads + reference arm), ncu launch list and full captures.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_balmer.log 2>&1; echo "rc=$?" >> gpurun_out/bench_balmer.log
timeout 600 python bench.py --workload multi_species --steps 20 --warmup 3 > gpurun_out/bench_multi.log 2>&1; echo "rc=$?" >> gpurun_out/bench_multi.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ref.log
tail -2 gpurun_out/bench_balmer.log | cut -c1-300; tail -2 gpurun_out/bench_multi.log | cut -c1-300; tail -2 gpurun_out/bench_ref.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_raw.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'los_stage|chord_stage|pix_contract|pixel_fold' --launch-skip 12 -c 4 -o gpurun_out/balmer_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
