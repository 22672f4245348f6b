#!/bin/bash
# compute-sanitizer passes over every kernel of the path ("race detection / sanitizers").
#   bash tools/sanitize.sh [outdir]        # on a GPU box; logs -> <outdir>/sanitize_<tool>.log (default gpurun_out)
# memcheck: out-of-bounds / misaligned accesses; racecheck: shared-memory hazards (the mbarrier pipelines of the two
# tcgen05 contractions, the block reductions of the chord / pixel kernels); synccheck: barrier misuse; initcheck:
# reads of uninitialised global memory (the workspaces).  Every pass is bounded by `timeout`.
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
CS=/usr/local/cuda/bin/compute-sanitizer
run() {  # tool, seconds, parts...
  local tool=$1 secs=$2; shift 2
  echo "== $tool: $*" | tee "$OUT/sanitize_$tool.log"
  timeout "$secs" $CS --tool "$tool" --print-limit 20 --error-exitcode 7 python tools/sanitize_target.py "$@" >> "$OUT/sanitize_$tool.log" 2>&1
  echo "exit code $?" | tee -a "$OUT/sanitize_$tool.log"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^OK|exit code" "$OUT/sanitize_$tool.log" | tail -12
}
run memcheck 420 balmer toeplitz multi hot ensemble hook
run synccheck 300 balmer toeplitz multi hook
run racecheck 600 balmer multi hook
run initcheck 300 balmer multi
