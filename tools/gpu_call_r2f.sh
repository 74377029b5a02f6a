#!/bin/bash
# Round 2, call F (N GPUs): default bench line on 1 GPU, then under torchrun on all visible GPUs.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2f
mkdir -p "$OUT"
N=$(nvidia-smi -L | wc -l)
leg() { local name=$1 secs=$2; shift 2; local t0=$(date +%s); timeout "$secs" "$@" > "$OUT/$name.log" 2>&1; echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a "$OUT/summary.txt"; }
leg pytest 600 python -m pytest tests/test_zzz_gpu_active_loop5.py tests/test_gpu_multirank.py -m gpu -q -p no:cacheprovider
leg bench_1 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e
if [ "$N" -gt 1 ]; then
  leg bench_$N 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-e2e
fi
cat "$OUT/summary.txt"
