#!/bin/bash
# GPU call: ncu launch list and full captures of the four hot-path kernels (balmer_series bench command).
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:chord_stage_kernel|if_contract_kernel|los_stage_kernel|pixel_stage_kernel' -s 12 -c 4 -f -o gpurun_out/prof_r1s2 $CMD > gpurun_out/ncu_full.log 2>&1
CMD2="python bench.py --workload multi_species --steps 3 --warmup 3 --no-cpu-baseline --batch 16384"
$CMD2 > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:chord_stage_kernel|los_stage_kernel' -s 6 -c 2 -f -o gpurun_out/prof_r1s2_multi $CMD2 > gpurun_out/ncu_full_multi.log 2>&1
tail -3 gpurun_out/ncu_full.log gpurun_out/ncu_full_multi.log
