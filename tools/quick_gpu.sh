#!/bin/bash
# Quick GPU iteration: parity tests of the posterior path, then short bench lines of the two main workloads for a list
# of model-option variants.   bash tools/quick_gpu.sh "opt1" "opt2" ...   ("" = defaults)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
if [ $# -eq 0 ]; then set -- ""; fi
for o in "$@"; do
  for w in ${WORKLOADS:-balmer_series multi_species}; do
    timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-extra --options "$o" > "gpurun_out/q_${w}_${o}.log" 2>gpurun_out/q.err || tail -3 gpurun_out/q.err
    echo "== $w [$o]"; python tools/bench_summary.py "gpurun_out/q_${w}_${o}.log" 2>/dev/null | head -3
  done
done
