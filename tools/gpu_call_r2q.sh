#!/bin/bash
# Round 2, call Q (2 GPUs): the two-rank NCCL worker on its own, with its stage reports.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2q
mkdir -p "$OUT"
leg() { local name=$1 secs=$2; shift 2; local t0=$(date +%s); timeout -k 5 "$secs" "$@" > "$OUT/$name.log" 2>&1; echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a "$OUT/summary.txt"; }
leg nccl_worker 220 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/nccl_worker.py
leg pytest_nccl 320 python -m pytest tests/test_gpu_nccl_two_ranks.py -m gpu -q -p no:cacheprovider
cat "$OUT/summary.txt"
