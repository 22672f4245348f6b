#!/bin/bash
# GPU call: ncu full capture of the folded contraction at bench size (TN = 128 launch of tools/pxc_probe.py)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pix_contract --launch-skip 1 -c 1 -o gpurun_out/pxc_full python tools/pxc_probe.py > gpurun_out/ncu_pxc.log 2>&1
tail -3 gpurun_out/ncu_pxc.log; ls -la gpurun_out/*.ncu-rep
