#!/bin/bash
# Round 2, final 1-GPU lease: every -m gpu test, smoke(), the full bench line (CPU baseline + end-to-end leg), the reference
# arm, the launch list of a step and ncu --set full of the four dominant kernels (summarised into profiles/ by
# tools/summarize_profiles.py).  ncu legs run only after the plain bench has exited 0.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r2z}
mkdir -p "$OUT"
leg() { local name=$1 secs=$2; shift 2; local t0=$(date +%s); timeout -k 5 "$secs" "$@" > "$OUT/$name.log" 2>&1; echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a "$OUT/summary.txt"; }
leg pytest_gpu 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -x
leg smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
leg bench 600 python bench.py
leg bench_reference 600 python bench.py --impl reference
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
leg ncu_launches 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file "$OUT/launches.csv" $B
for k in k_forcing_v7 k_stress_v7 k1_mark2 k5c_run k7_prepare; do
  leg ncu_$k 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o "$OUT/$k" -f $B
done
cat "$OUT/summary.txt"
