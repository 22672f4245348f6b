#!/bin/bash
# GPU call: ncu full capture of the spectral and LOS kernels (balmer_series) after the session-2 optimisations.
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:chord_stage_kernel|los_stage_kernel' -s 6 -c 2 -f -o gpurun_out/prof_r1s2c $CMD > gpurun_out/ncu_full.log 2>&1
CMD2="python bench.py --workload multi_species --steps 3 --warmup 3 --no-cpu-baseline --batch 16384"
$CMD2 > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:chord_stage_kernel|los_stage_kernel' -s 6 -c 2 -f -o gpurun_out/prof_r1s2c_multi $CMD2 > gpurun_out/ncu_full_multi.log 2>&1
tail -n 3 gpurun_out/ncu_full.log gpurun_out/ncu_full_multi.log
