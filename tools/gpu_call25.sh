#!/bin/bash
# GPU call: ncu full capture of the multi-species step (LOS + chord stage).
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'los_stage|chord_stage' --launch-skip 6 -c 2 -o gpurun_out/multi_full python bench.py --workload multi_species --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_multi.log 2>&1
ls -la gpurun_out/multi_full.ncu-rep
