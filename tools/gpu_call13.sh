#!/bin/bash
# GPU call: per-role wait-cycle trace of the instrument contraction and its debug modes.
mkdir -p gpurun_out
for dbg in 0 1 4 8 16 28; do
  echo "=== BAYSAR_IFC_DEBUG=$dbg" >> gpurun_out/ifc_trace.log
  BAYSAR_IFC_TRACE=1 BAYSAR_IFC_DEBUG=$dbg timeout 120 python tools/ifc_probe.py 4096 8002 >> gpurun_out/ifc_trace.log 2>&1
done
cat gpurun_out/ifc_trace.log
