#!/bin/bash
# GPU call (2 GPUs): smoke(), new reproducibility test, ensemble bench for config 4 (32 chords) on 1 and 2 GPUs.
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout 300 python -m pytest tests/test_gpu_pix_contract.py -m gpu -x -q > gpurun_out/pytest_pxc.log 2>&1; tail -2 gpurun_out/pytest_pxc.log
CUDA_VISIBLE_DEVICES=0 timeout 900 python tools/ensemble_bench.py --walkers-per-gpu 32768 --steps 2 > gpurun_out/ensemble_n1.log 2>&1; echo "rc=$?" >> gpurun_out/ensemble_n1.log; tail -2 gpurun_out/ensemble_n1.log | cut -c1-500
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/ensemble_bench.py --walkers-per-gpu 32768 --steps 2 > gpurun_out/ensemble_n2.log 2>&1; echo "rc=$?" >> gpurun_out/ensemble_n2.log; tail -2 gpurun_out/ensemble_n2.log | cut -c1-500
