#!/bin/bash
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2g
mkdir -p "$OUT"
leg() { local name=$1 secs=$2; shift 2; local t0=$(date +%s); timeout "$secs" "$@" > "$OUT/$name.log" 2>&1; echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a "$OUT/summary.txt"; }
leg ncu_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file "$OUT/launches.csv" python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e
cat "$OUT/summary.txt"
