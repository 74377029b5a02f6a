#!/bin/bash
# Round 2, call A: seventh-generation kernels on hardware — parity, bench line, launch list, one full capture each.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2a
mkdir -p "$OUT"
leg() { local name=$1 secs=$2; shift 2; local t0=$(date +%s); timeout "$secs" "$@" > "$OUT/$name.log" 2>&1; echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a "$OUT/summary.txt"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > "$OUT/gpu.txt" 2>&1
leg pytest_v7 600 python -m pytest tests/test_gpu_v7_kernels.py tests/test_gpu_forcing.py tests/test_gpu_stress.py -m gpu -q -p no:cacheprovider -x
leg bench_v7 600 python bench.py --steps 5 --warmup 3 --variant 7 --no-cpu-baseline --no-e2e
leg bench_v2 600 python bench.py --steps 5 --warmup 3 --variant 2 --no-cpu-baseline --no-e2e
leg ncu_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file "$OUT/launches_v7.csv" python bench.py --steps 1 --warmup 3 --variant 7 --no-cpu-baseline --no-e2e
leg ncu_forcing_v7 900 ncu --set full --clock-control none --import-source on -k regex:k_forcing_v7 -s 3 -c 1 -o "$OUT/forcing_v7" -f python bench.py --steps 1 --warmup 3 --variant 7 --no-cpu-baseline --no-e2e
leg ncu_stress_v7 900 ncu --set full --clock-control none --import-source on -k regex:k_stress_v7 -s 3 -c 1 -o "$OUT/stress_v7" -f python bench.py --steps 1 --warmup 3 --variant 7 --no-cpu-baseline --no-e2e
cat "$OUT/summary.txt"
