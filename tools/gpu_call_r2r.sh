#!/bin/bash
# Round 2, call R (N GPUs): owned loop 5 over NCCL - parity worker, bench at 1..N GPUs, owned against replicated.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2r
mkdir -p "$OUT"
N=$(nvidia-smi -L | wc -l)
leg() { local name=$1 secs=$2; shift 2; local t0=$(date +%s); timeout -k 5 "$secs" "$@" > "$OUT/$name.log" 2>&1; echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a "$OUT/summary.txt"; }
leg nccl_worker 220 env IBM_DEBUG_OWNED=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/nccl_worker.py
leg pytest_nccl 320 python -m pytest tests/test_gpu_nccl_two_ranks.py tests/test_gpu_multirank.py -m gpu -q -p no:cacheprovider
leg bench_1 300 python bench.py --no-cpu-baseline --no-e2e
for n in 2 4 8; do
  if [ "$N" -ge "$n" ]; then
    leg bench_$n 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --no-cpu-baseline --no-e2e
    leg bench_${n}_replicated 300 env IBM_OWNED_LOOP5=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --no-cpu-baseline --no-e2e
  fi
done
cat "$OUT/summary.txt"
