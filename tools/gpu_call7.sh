#!/bin/bash
# GPU call: per-role wait-cycle trace of the contraction kernel + MMA-mix timing experiments.
mkdir -p gpurun_out
rm -f gpurun_out/ifc_trace.log
for d in 0 1 28 29; do
  echo "== BAYSAR_IFC_DEBUG=$d" >> gpurun_out/ifc_trace.log
  BAYSAR_IFC_TRACE=1 BAYSAR_IFC_DEBUG=$d timeout 120 python tools/ifc_probe.py 4096 8002 2>&1 | grep -E "kernel|trace" | cut -c1-100 | tail -9 >> gpurun_out/ifc_trace.log
done
cat gpurun_out/ifc_trace.log
