#!/bin/bash
# Round 2, call S (1 GPU): launch records of the staged kernels (k7_prepare) - parity, bench, launch list, ncu of the two kernels.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r2s}
mkdir -p "$OUT"
leg() { local name=$1 secs=$2; shift 2; local t0=$(date +%s); timeout -k 5 "$secs" "$@" > "$OUT/$name.log" 2>&1; echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a "$OUT/summary.txt"; }
leg pytest 400 python -m pytest tests/test_gpu_v7_kernels.py tests/test_gpu_forcing.py tests/test_gpu_stress.py tests/test_gpu_marking.py tests/test_gpu_zz_reference_vectors.py tests/test_gpu_multirank.py -m gpu -q -p no:cacheprovider -x
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
leg bench 300 python bench.py --no-cpu-baseline --no-e2e
leg ncu_launches 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file "$OUT/launches.csv" $B
for k in k_forcing_v7 k_stress_v7; do
  leg ncu_$k 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o "$OUT/$k" -f $B
done
cat "$OUT/summary.txt"
