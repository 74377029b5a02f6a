#!/bin/bash
# Round 2, call T (1 GPU): end-to-end leg - threads / chunk size of the staged host copies; latitude order of the stress points.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r2t}
mkdir -p "$OUT"
leg() { local name=$1 secs=$2; shift 2; local t0=$(date +%s); timeout -k 5 "$secs" "$@" > "$OUT/$name.log" 2>&1; echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a "$OUT/summary.txt"; }
nproc > "$OUT/host.txt"; lscpu | head -25 >> "$OUT/host.txt"; free -g >> "$OUT/host.txt"
leg pytest 300 python -m pytest tests/test_gpu_v7_kernels.py tests/test_gpu_stress.py -m gpu -q -p no:cacheprovider -x
leg e2e_default 300 python bench.py --no-cpu-baseline --steps 5
leg e2e_t16 300 env IBM_COPY_THREADS=16 python bench.py --no-cpu-baseline --steps 3
leg e2e_t4 300 env IBM_COPY_THREADS=4 python bench.py --no-cpu-baseline --steps 3
leg e2e_t16_c16 300 env IBM_COPY_THREADS=16 IBM_CHUNK_MB=16 python bench.py --no-cpu-baseline --steps 3
leg e2e_t12_c256 300 env IBM_COPY_THREADS=12 IBM_CHUNK_MB=256 python bench.py --no-cpu-baseline --steps 3
cat "$OUT/summary.txt"
