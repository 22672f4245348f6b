#!/bin/bash
# GPU call: per-role wait-cycle trace of the contraction kernel.
mkdir -p gpurun_out
for d in 0 28; do
  echo "== BAYSAR_IFC_DEBUG=$d" >> gpurun_out/ifc_trace.log
  BAYSAR_IFC_TRACE=1 BAYSAR_IFC_DEBUG=$d timeout 120 python tools/ifc_probe.py 4096 8002 2>&1 | grep -E "kernel|trace" | cut -c1-100 | tail -9 >> gpurun_out/ifc_trace.log
done
cat gpurun_out/ifc_trace.log
