#!/bin/bash
# Round 2, call AF (2 GPUs): the driver's scaling leg as it launches it - default flags (end-to-end leg included) on 2 ranks.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r2af}
mkdir -p "$OUT"
leg() { local name=$1 secs=$2; shift 2; local t0=$(date +%s); timeout -k 5 "$secs" "$@" > "$OUT/$name.log" 2>&1; echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a "$OUT/summary.txt"; }
leg bench_2_default 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3
leg bench_2_reference 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 3 --warmup 1
cat "$OUT/summary.txt"
