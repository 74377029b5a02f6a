#!/bin/bash
# Round 2, call H: the whole -m gpu suite, smoke(), the default bench line (with e2e and CPU baselines) and the reference arm.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2h
mkdir -p "$OUT"
leg() { local name=$1 secs=$2; shift 2; local t0=$(date +%s); timeout "$secs" "$@" > "$OUT/$name.log" 2>&1; echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a "$OUT/summary.txt"; }
leg pytest_gpu 1500 python -m pytest tests -m gpu -q -p no:cacheprovider
leg smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
leg bench 900 python bench.py
leg bench_reference 600 python bench.py --impl reference --steps 2 --warmup 1
cat "$OUT/summary.txt"
