#!/bin/bash
# Round 2, last lease (1 GPU): every -m gpu test and smoke() on the final tree.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r2last}
mkdir -p "$OUT"
leg() { local name=$1 secs=$2; shift 2; local t0=$(date +%s); timeout -k 5 "$secs" "$@" > "$OUT/$name.log" 2>&1; echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a "$OUT/summary.txt"; }
leg pytest_gpu 600 python -m pytest tests -m gpu -q -p no:cacheprovider -x
leg smoke 100 python -c "import __graft_entry__ as g; g.smoke()"
cat "$OUT/summary.txt"
