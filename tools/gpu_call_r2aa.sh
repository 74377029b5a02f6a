#!/bin/bash
# Round 2, call AA (1 GPU): schedule order + launch records beside the marking kernel - parity, timeline, bench.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r2aa}
mkdir -p "$OUT"
leg() { local name=$1 secs=$2; shift 2; local t0=$(date +%s); timeout -k 5 "$secs" "$@" > "$OUT/$name.log" 2>&1; echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a "$OUT/summary.txt"; }
leg pytest_gpu 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x
leg bench_timing 300 env IBM_DEBUG_TIMING=1 python bench.py --no-cpu-baseline --no-e2e
leg bench 300 python bench.py --no-cpu-baseline --no-e2e
cat "$OUT/summary.txt"
