#!/bin/bash
# Evidence runs, one part per gpurun call (what a call writes under gpurun_out/ must stay below 64 MiB):
#   bash tools/gpu_evidence.sh a   -> GPU tests, default bench, reference arm, launch list, ncu --set full of the Balmer step
#   bash tools/gpu_evidence.sh b   -> ncu --set full of the multi-species LOS and chord kernels
#   bash tools/gpu_evidence.sh c   -> ncu --set full of the hot-Balmer chord kernel
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
COMMON="--steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-extra"
case ${1:-a} in
  a)
    bash tools/gpu_round.sh tests bench ref
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_raw.csv python bench.py $COMMON > gpurun_out/ncu_launch.log 2>&1
    timeout 900 $NCU -k regex:'los_stage|chord_stage|pix_contract|pixel_fold' --launch-skip 12 -c 4 -o gpurun_out/balmer_full -f python bench.py $COMMON > gpurun_out/ncu_full.log 2>&1 ;;
  b)
    timeout 900 $NCU -k regex:'los_stage|chord_stage' --launch-skip 6 -c 2 -o gpurun_out/multi_full -f python bench.py --workload multi_species $COMMON > gpurun_out/ncu_full_multi.log 2>&1 ;;
  c)
    timeout 900 $NCU -k regex:'chord_stage' --launch-skip 3 -c 1 -o gpurun_out/hot_full -f python bench.py --workload balmer_hot $COMMON > gpurun_out/ncu_full_hot.log 2>&1 ;;
esac
ls -la gpurun_out/*.ncu-rep 2>/dev/null
