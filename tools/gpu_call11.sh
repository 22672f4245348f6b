#!/bin/bash
# GPU call: parity suite (incl. ensemble + gradient tests) and quick benches.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
bash tools/gpu_call10.sh
