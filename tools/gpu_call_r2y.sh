#!/bin/bash
# Round 2, call Y (1 GPU): L2 prefetch of the bounding rows on / off per kernel.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r2y}
mkdir -p "$OUT"
leg() { local name=$1 secs=$2; shift 2; local t0=$(date +%s); timeout -k 5 "$secs" "$@" > "$OUT/$name.log" 2>&1; echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a "$OUT/summary.txt"; }
for pf in 3 0 1 2; do
  leg bench_pf$pf 300 env IBM_V7_PREFETCH=$pf python bench.py --no-cpu-baseline --no-e2e --steps 6
done
cat "$OUT/summary.txt"
