#!/bin/bash
# GPU call: full evidence run after the LOS / component set-up changes: parity suite, bench lines (both workloads +
# reference arm), ncu launch list and full captures of both workloads.
bash tools/gpu_call19.sh
bash tools/gpu_call25.sh
