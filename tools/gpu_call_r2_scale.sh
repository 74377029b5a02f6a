#!/bin/bash
# Round 2: the scaling run the driver does at round end, on one N-GPU box: bench at 1, 2, 4, ... N ranks (device-resident leg).
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r2scale}
mkdir -p "$OUT"
N=$(nvidia-smi -L | wc -l)
leg() { local name=$1 secs=$2; shift 2; local t0=$(date +%s); timeout -k 5 "$secs" "$@" > "$OUT/$name.log" 2>&1; echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a "$OUT/summary.txt"; }
leg bench_1 200 python bench.py --no-cpu-baseline --no-e2e
for n in 2 4 8; do
  if [ "$N" -ge "$n" ]; then
    leg bench_$n 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --no-cpu-baseline --no-e2e
  fi
done
leg timing_$N 200 env IBM_DEBUG_TIMING=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus $N --no-cpu-baseline --no-e2e
cat "$OUT/summary.txt"
