#!/bin/bash
# Round 2, call AB (1 GPU): the slowest block of k5c_run against the partner-list length and the packing of components into blocks.
# (negative result, DESIGN.md 4.4: the switches / the four-lane kernels existed only in the working tree of this experiment)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r2ab}
mkdir -p "$OUT"
leg() { local name=$1 secs=$2; shift 2; local t0=$(date +%s); timeout -k 5 "$secs" "$@" > "$OUT/$name.log" 2>&1; echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a "$OUT/summary.txt"; }
leg pytest 600 python -m pytest tests/test_zzz_gpu_active_loop5.py tests/test_gpu_motion.py tests/test_gpu_zz_reference_vectors.py -m gpu -q -p no:cacheprovider -x
for cfg in "12 56" "6 56" "12 32" "12 20" "12 12"; do
  set -- $cfg
  leg bench_np$1_fill$2 200 env IBM_DEBUG_TIMING=1 IBM_K5C_NP=$1 IBM_K5C_FILL=$2 python bench.py --no-cpu-baseline --no-e2e --steps 6
done
cat "$OUT/summary.txt"
