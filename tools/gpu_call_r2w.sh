#!/bin/bash
# Round 2, call W (1 GPU): threads per CTA of the staged kernels, timeline medians, prepare with early exit.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r2w}
mkdir -p "$OUT"
leg() { local name=$1 secs=$2; shift 2; local t0=$(date +%s); timeout -k 5 "$secs" "$@" > "$OUT/$name.log" 2>&1; echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a "$OUT/summary.txt"; }
leg pytest 300 python -m pytest tests/test_gpu_v7_kernels.py tests/test_gpu_stress.py tests/test_gpu_forcing.py -m gpu -q -p no:cacheprovider -x
leg bench_256 300 env IBM_DEBUG_TIMING=1 python bench.py --no-cpu-baseline --no-e2e
for nt in 192 224 320; do
  leg bench_$nt 300 env IBM_V7_THREADS=$nt python bench.py --no-cpu-baseline --no-e2e --steps 5
done
cat "$OUT/summary.txt"
