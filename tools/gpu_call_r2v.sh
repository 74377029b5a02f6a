#!/bin/bash
# Round 2, call V (N GPUs): timeline of a step on 1 and N ranks (IBM_DEBUG_TIMING), stress kernel capture at 1 GPU.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r2v}
mkdir -p "$OUT"
N=$(nvidia-smi -L | wc -l)
leg() { local name=$1 secs=$2; shift 2; local t0=$(date +%s); timeout -k 5 "$secs" "$@" > "$OUT/$name.log" 2>&1; echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a "$OUT/summary.txt"; }
leg timing_1 300 env IBM_DEBUG_TIMING=1 python bench.py --no-cpu-baseline --no-e2e
leg timing_$N 300 env IBM_DEBUG_TIMING=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --no-cpu-baseline --no-e2e
leg bench_$N 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --no-cpu-baseline --no-e2e
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
leg ncu_k_stress_v7 300 ncu --set full --clock-control none --import-source on -k regex:k_stress_v7 -s 3 -c 1 -o "$OUT/k_stress_v7" -f $B
cat "$OUT/summary.txt"
