#!/bin/bash
# GPU call: BASELINE configs 4 and 5: 32-chord bench line (1 GPU share of the walkers) and the shape sweep.
mkdir -p gpurun_out
timeout 900 python bench.py --workload multi_species_32chords --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_32chords.log 2>&1; echo "rc=$?" >> gpurun_out/bench_32chords.log
tail -2 gpurun_out/bench_32chords.log | cut -c1-400
timeout 1200 python tools/shape_sweep.py --out gpurun_out/shape_sweep.jsonl > gpurun_out/shape_sweep.log 2>&1; echo "rc=$?" >> gpurun_out/shape_sweep.log
tail -20 gpurun_out/shape_sweep.log | cut -c1-250
