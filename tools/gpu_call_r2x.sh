#!/bin/bash
# Round 2, call X (N GPUs): halo + owned loop 5 parity on hardware, timeline medians on N ranks, bench 1..N.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/${1:-r2x}
mkdir -p "$OUT"
N=$(nvidia-smi -L | wc -l)
leg() { local name=$1 secs=$2; shift 2; local t0=$(date +%s); timeout -k 5 "$secs" "$@" > "$OUT/$name.log" 2>&1; echo "$name rc=$? $(( $(date +%s) - t0 ))s" | tee -a "$OUT/summary.txt"; }
leg pytest 400 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_nccl_two_ranks.py tests/test_gpu_marking.py tests/test_gpu_stress.py tests/test_gpu_v7_kernels.py -m gpu -q -p no:cacheprovider -x
leg bench_1 300 python bench.py --no-cpu-baseline --no-e2e
leg timing_$N 300 env IBM_DEBUG_TIMING=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --no-cpu-baseline --no-e2e
for n in 2 4 8; do
  if [ "$N" -ge "$n" ]; then
    leg bench_$n 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --no-cpu-baseline --no-e2e
  fi
done
cat "$OUT/summary.txt"
