#!/bin/bash
# Round 2, call I (N GPUs): NCCL two-rank parity test, bench at 1, 2, ... N GPUs, marking kernel variants.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2i
mkdir -p "$OUT"
N=$(nvidia-smi -L | wc -l)
leg() { local name=$1 secs=$2; shift 2; local t0=$(date +%s); timeout "$secs" "$@" > "$OUT/$name.log" 2>&1; echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a "$OUT/summary.txt"; }
leg pytest_nccl 600 python -m pytest tests/test_gpu_nccl_two_ranks.py tests/test_gpu_multirank.py tests/test_zz_gpu_shim_dropin.py -m gpu -q -p no:cacheprovider
leg bench_1 600 python bench.py --no-cpu-baseline --no-e2e
leg bench_1_mark2 600 env IBM_MARK_VARIANT=1 python bench.py --no-cpu-baseline --no-e2e
for n in 2 4 8; do
  if [ "$N" -ge "$n" ]; then
    leg bench_$n 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --no-cpu-baseline --no-e2e
  fi
done
cat "$OUT/summary.txt"
