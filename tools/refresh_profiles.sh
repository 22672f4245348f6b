#!/bin/bash
# Turn the artefacts of tools/gpu_evidence.sh a|b|c (gpurun_out/) into the tracked summaries under profiles/.
C="ncu --set full --clock-control none --import-source on"; S="--steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-extra"
python tools/ncu_counters.py gpurun_out/balmer_full.ncu-rep --workload balmer_series --batch 4096 --out profiles/r2_balmer_series_counters.json --command "$C -k regex:'los_stage|chord_stage|pix_contract|pixel_fold' --launch-skip 12 -c 4 python bench.py $S"
python tools/ncu_counters.py gpurun_out/multi_full.ncu-rep --workload multi_species --batch 65536 --out profiles/r2_multi_species_counters.json --command "$C -k regex:'los_stage|chord_stage' --launch-skip 6 -c 2 python bench.py --workload multi_species $S"
python tools/ncu_counters.py gpurun_out/hot_full.ncu-rep --workload balmer_hot --batch 4096 --out profiles/r2_balmer_hot_counters.json --command "$C -k regex:'chord_stage' --launch-skip 3 -c 1 python bench.py --workload balmer_hot $S"
python tools/ncu_hotspots.py gpurun_out/balmer_full.ncu-rep "" 25 > profiles/r2_final_balmer_series_ncu_full.txt 2>&1
python tools/ncu_hotspots.py gpurun_out/multi_full.ncu-rep "" 25 > profiles/r2_final_multi_species_ncu_full.txt 2>&1
python tools/ncu_hotspots.py gpurun_out/hot_full.ncu-rep "" 25 > profiles/r2_final_balmer_hot_ncu_full.txt 2>&1
cp gpurun_out/bench_default.log profiles/r2_final_bench_default.json; cp gpurun_out/bench_ref.log profiles/r2_final_bench_ref.json
cp gpurun_out/pytest_gpu.log profiles/r2_final_pytest_gpu.log; cp gpurun_out/launches_raw.csv profiles/r2_final_launches_raw.csv
