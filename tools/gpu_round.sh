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
      echo 'ncu captures: tools/gpu_evidence.sh (one capture per gpurun call, 64 MiB limit)' ;;
  esac
done
