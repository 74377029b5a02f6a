#!/bin/bash
# Round 2, call D: loop 5 cluster mode with resident state + cheaper classification; stress with 3 CTAs/SM.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2d
mkdir -p "$OUT"
leg() { local name=$1 secs=$2; shift 2; local t0=$(date +%s); timeout "$secs" "$@" > "$OUT/$name.log" 2>&1; echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a "$OUT/summary.txt"; }
leg pytest_loop5 900 python -m pytest tests/test_zzz_gpu_active_loop5.py tests/test_gpu_motion.py tests/test_gpu_zz_reference_vectors.py tests/test_gpu_v7_kernels.py -m gpu -q -p no:cacheprovider
leg bench_split 600 python bench.py --steps 5 --warmup 3 --variant 7 --split 1 --no-cpu-baseline --no-e2e
leg bench_split_224 600 env IBM_V7_THREADS=224 python bench.py --steps 5 --warmup 3 --variant 7 --split 1 --no-cpu-baseline --no-e2e
leg ncu_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$OUT/launches.csv" python bench.py --steps 1 --warmup 3 --variant 7 --split 1 --no-cpu-baseline --no-e2e
leg ncu_stress_v7 900 ncu --set full --clock-control none --import-source on -k regex:k_stress_v7 -s 3 -c 1 -o "$OUT/stress_v7" -f python bench.py --steps 1 --warmup 3 --variant 7 --split 1 --no-cpu-baseline --no-e2e
cat "$OUT/summary.txt"
