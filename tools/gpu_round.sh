#!/bin/bash
# One GPU evidence run: parity suite, bench lines (both arms), optional sanitizer passes.  Logs -> gpurun_out/.
#   bash tools/gpu_round.sh [tests] [bench] [ref] [sanitize] [ncu]
mkdir -p gpurun_out
WHAT=${*:-tests bench ref}
for w in $WHAT; do
  case $w in
    tests)
      timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
      tail -15 gpurun_out/pytest_gpu.log ;;
    bench)
      timeout 900 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "rc=$?" >> gpurun_out/bench_default.err
      tail -3 gpurun_out/bench_default.err; python tools/bench_summary.py gpurun_out/bench_default.log ;;
    ref)
      timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-200 ;;
    sanitize)
      bash tools/sanitize.sh gpurun_out ;;
    ncu)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_raw.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-extra > gpurun_out/ncu_launch.log 2>&1
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:'los_stage|chord_stage|pix_contract|pixel_fold' --launch-skip 12 -c 4 -o gpurun_out/balmer_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-extra > gpurun_out/ncu_full.log 2>&1
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:'los_stage|chord_stage' --launch-skip 6 -c 2 -o gpurun_out/multi_full -f python bench.py --workload multi_species --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-extra > gpurun_out/ncu_full_multi.log 2>&1
      ls -la gpurun_out/*.ncu-rep ;;
  esac
done
