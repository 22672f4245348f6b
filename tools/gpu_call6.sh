#!/bin/bash
# GPU call: ncu full capture of the contraction kernel alone (tools/ifc_probe.py, 4096 x 8002, 431-tap band).
mkdir -p gpurun_out
CMD="python tools/ifc_probe.py 4096 8002"
$CMD > gpurun_out/plain_ifc.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:if_contract_kernel -s 2 -c 1 -f -o gpurun_out/prof_ifc2 $CMD > gpurun_out/ncu_ifc2.log 2>&1
tail -n 3 gpurun_out/plain_ifc.log gpurun_out/ncu_ifc2.log
