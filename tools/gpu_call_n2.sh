#!/bin/bash
# 2-GPU call: weak-scaling bench lines (torchrun, NCCL) for both workloads.
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_n2_balmer.log 2>&1; echo "rc=$?" >> gpurun_out/bench_n2_balmer.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --workload multi_species > gpurun_out/bench_n2_multi.log 2>&1; echo "rc=$?" >> gpurun_out/bench_n2_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --impl reference > gpurun_out/bench_n2_ref.log 2>&1; echo "rc=$?" >> gpurun_out/bench_n2_ref.log
tail -2 gpurun_out/bench_n2_balmer.log | cut -c1-400; tail -2 gpurun_out/bench_n2_multi.log | cut -c1-400; tail -2 gpurun_out/bench_n2_ref.log | cut -c1-400
