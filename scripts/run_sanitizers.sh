#!/bin/bash
# compute-sanitizer evidence for the hand-written kernels: memcheck over every kernel family,
# racecheck (shared-memory hazards) on K4's lock-free union-find, the pipelined K1 and the tensor-core kernels.
# Usage (on the GPU box): bash scripts/run_sanitizers.sh [outdir]   -> <outdir>/sanitizer_*.log
out=${1:-gpurun_out}
mkdir -p "$out"
CS=/usr/local/cuda/bin/compute-sanitizer
run() {  # name tool families...
  name=$1; tool=$2; shift 2
  timeout 900 $CS --tool "$tool" --print-limit 20 --error-exitcode 0 python scripts/sanitize_driver.py "$@" \
      > "$out/sanitizer_${tool}_${name}.log" 2>&1
  echo "== $tool $name: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$out/sanitizer_${tool}_${name}.log" | tail -1)"
}
run all memcheck
run k4 racecheck k4
run k1 racecheck k1
run k1t racecheck k1t
run k2k3 racecheck k2 k3
run k2w racecheck k2w
