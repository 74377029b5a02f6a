#!/bin/bash
# Round 2: the ncu evidence of the shipped defaults - launch list of one step and --set full captures of the dominant kernels.
# Run only after `python bench.py` has exited 0 without ncu (leg bench below); outputs are summarised into profiles/ by
# tools/summarize_profiles.py.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2p
mkdir -p "$OUT"
leg() { local name=$1 secs=$2; shift 2; local t0=$(date +%s); timeout "$secs" "$@" > "$OUT/$name.log" 2>&1; echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a "$OUT/summary.txt"; }
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
leg bench 300 python bench.py --no-cpu-baseline --no-e2e
leg ncu_launches 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file "$OUT/launches.csv" $B
for k in k_forcing_v7 k_stress_v7 k1_mark2 k5c_run; do
  leg ncu_$k 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o "$OUT/$k" -f $B
done
cat "$OUT/summary.txt"
