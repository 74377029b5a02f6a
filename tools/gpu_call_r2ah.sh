#!/bin/bash
# Round 2, call AH (1 GPU): lattice index split without integer divisions in the forcing kernel - parity at the BASELINE sizes, bench.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r2ah}
mkdir -p "$OUT"
leg() { local name=$1 secs=$2; shift 2; local t0=$(date +%s); timeout -k 5 "$secs" "$@" > "$OUT/$name.log" 2>&1; echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a "$OUT/summary.txt"; }
leg pytest 400 python -m pytest tests/test_gpu_v7_kernels.py tests/test_gpu_forcing.py tests/test_gpu_zz_baseline_sizes.py tests/test_gpu_zz_reference_vectors.py tests/test_zz_gpu_shim_dropin.py tests/test_gpu_motion.py tests/test_gpu_zz_device_api.py -m gpu -q -p no:cacheprovider -x
leg smoke 100 python -c "import __graft_entry__ as g; g.smoke()"
leg bench 300 python bench.py --no-cpu-baseline --no-e2e
cat "$OUT/summary.txt"
