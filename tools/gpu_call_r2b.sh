#!/bin/bash
# Round 2, call B: v7 after the instruction diet — parity on hardware, block sizes, one full capture.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2b
mkdir -p "$OUT"
leg() { local name=$1 secs=$2; shift 2; local t0=$(date +%s); timeout "$secs" "$@" > "$OUT/$name.log" 2>&1; echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a "$OUT/summary.txt"; }
leg pytest_v7 600 python -m pytest tests/test_gpu_v7_kernels.py tests/test_gpu_forcing.py tests/test_gpu_stress.py -m gpu -q -p no:cacheprovider -x
for nt in 256 320 192; do
  leg bench_v7_$nt 600 env IBM_V7_THREADS=$nt python bench.py --steps 5 --warmup 3 --variant 7 --no-cpu-baseline --no-e2e
done
leg ncu_forcing_v7 900 env IBM_V7_THREADS=320 ncu --set full --clock-control none --import-source on -k regex:k_forcing_v7 -s 3 -c 1 -o "$OUT/forcing_v7" -f python bench.py --steps 1 --warmup 3 --variant 7 --no-cpu-baseline --no-e2e
leg ncu_stress_v7 900 env IBM_V7_THREADS=320 ncu --set full --clock-control none --import-source on -k regex:k_stress_v7 -s 3 -c 1 -o "$OUT/stress_v7" -f python bench.py --steps 1 --warmup 3 --variant 7 --no-cpu-baseline --no-e2e
cat "$OUT/summary.txt"
