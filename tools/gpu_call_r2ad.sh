#!/bin/bash
# Round 2, call AD (N GPUs): the forcing field re-zeroed by the stress kernel on the way - every GPU test, bench 1 and N.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r2ad}
mkdir -p "$OUT"
N=$(nvidia-smi -L | wc -l)
leg() { local name=$1 secs=$2; shift 2; local t0=$(date +%s); timeout -k 5 "$secs" "$@" > "$OUT/$name.log" 2>&1; echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a "$OUT/summary.txt"; }
leg pytest_gpu 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x
leg bench_1 300 env IBM_DEBUG_TIMING=1 python bench.py --no-cpu-baseline --no-e2e
if [ "$N" -ge 2 ]; then
  leg bench_$N 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --no-cpu-baseline --no-e2e
fi
cat "$OUT/summary.txt"
