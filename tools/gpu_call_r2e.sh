#!/bin/bash
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2e
mkdir -p "$OUT"
leg() { local name=$1 secs=$2; shift 2; local t0=$(date +%s); timeout "$secs" "$@" > "$OUT/$name.log" 2>&1; echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a "$OUT/summary.txt"; }
leg pytest 900 python -m pytest tests/test_zzz_gpu_active_loop5.py tests/test_gpu_v7_kernels.py -m gpu -q -p no:cacheprovider -k "cluster or v7"
leg bench_split 600 python bench.py --steps 5 --warmup 3 --variant 7 --split 1 --no-cpu-baseline --no-e2e
leg ncu_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$OUT/launches.csv" python bench.py --steps 1 --warmup 3 --variant 7 --split 1 --no-cpu-baseline --no-e2e
cat "$OUT/summary.txt"
