#!/bin/bash
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2j
mkdir -p "$OUT"
leg() { local name=$1 secs=$2; shift 2; local t0=$(date +%s); timeout "$secs" "$@" > "$OUT/$name.log" 2>&1; echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a "$OUT/summary.txt"; }
leg pytest 300 python -m pytest tests/test_gpu_multirank.py tests/test_zzz_gpu_active_loop5.py tests/test_zzz_gpu_mark_variant.py tests/test_gpu_v7_kernels.py -m gpu -q -p no:cacheprovider -x
leg bench_1 300 python bench.py --no-cpu-baseline --no-e2e
leg bench_1_soa 300 env IBM_V7_STRESS_SOA=1 python bench.py --no-cpu-baseline --no-e2e --steps 5
leg ncu_launches 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file "$OUT/launches.csv" python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e
cat "$OUT/summary.txt"
