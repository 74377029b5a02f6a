#!/bin/bash
# Round 2, call AC (1 GPU): four lanes per instance in the budget scans (k5q_nbrmax) - parity, launch list, bench.
# (negative result, DESIGN.md 4.4: the switches / the four-lane kernels existed only in the working tree of this experiment)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r2ac}
mkdir -p "$OUT"
leg() { local name=$1 secs=$2; shift 2; local t0=$(date +%s); timeout -k 5 "$secs" "$@" > "$OUT/$name.log" 2>&1; echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a "$OUT/summary.txt"; }
leg pytest 600 python -m pytest tests/test_zzz_gpu_active_loop5.py tests/test_gpu_motion.py tests/test_gpu_zz_reference_vectors.py tests/test_gpu_multirank.py -m gpu -q -p no:cacheprovider -x
leg bench 300 python bench.py --no-cpu-baseline --no-e2e
leg ncu_launches 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file "$OUT/launches.csv" python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e
cat "$OUT/summary.txt"
