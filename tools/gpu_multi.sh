#!/bin/bash
# Multi-GPU evidence: the 2-rank GPU tests and the bench lines at N ranks the way the driver launches them.
#   gpurun --gpus N -- 'bash tools/gpu_multi.sh N'
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ensemble.py tests/test_gpu_parity.py -m gpu -x -q -k "two_gpu or 2gpu or global or full_batch" > gpurun_out/pytest_gpu_n$N.log 2>&1; tail -3 gpurun_out/pytest_gpu_n$N.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "rc=$?"
tail -2 gpurun_out/bench_n$N.err; python tools/bench_summary.py gpurun_out/bench_n$N.log | grep -v "^ *+* *[a-z]* *[0-9.]* ms fp32"
