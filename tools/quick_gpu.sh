mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
for o in "" "chord_min_blocks=4" "chord_min_blocks=3"; do
  for w in balmer_series multi_species; do
    timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-extra --options "$o" > gpurun_out/q_${w}_${o}.log 2>gpurun_out/q.err || tail -3 gpurun_out/q.err
    echo "== $w [$o]"; python tools/bench_summary.py gpurun_out/q_${w}_${o}.log | head -4
  done
done
