#!/bin/bash
# First GPU call of a round: everything that was written without hardware gets measured in ONE gpurun lease.
#   gpurun --timeout 2400 -- 'bash tools/gpu_first_call.sh'
# Outputs land in gpurun_out/first_call/ (merged back by gpurun); nothing here reads /root/reference.
# Every leg runs under its own `timeout`, so a hanging kernel costs that leg, not the lease; legs are independent
# (a failure is recorded and the script goes on).
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/first_call
mkdir -p "$OUT"
PY=python
leg() {  # leg <name> <seconds> <command...>
  local name=$1 secs=$2; shift 2
  local t0=$(date +%s)
  timeout "$secs" "$@" > "$OUT/$name.log" 2>&1
  local rc=$?
  echo "$name rc=$rc $(( $(date +%s) - t0 ))s" | tee -a "$OUT/summary.txt"
}

nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > "$OUT/gpu.txt" 2>&1

# 0. secondary ceilings of this box: fp64 FMA / rsqrt rate, launch latency (stream, graph), grid-barrier latency, copy bandwidth
[ -x tools/microbench/microbench ] || nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench/microbench tools/microbench/microbench.cu
leg microbench 180 tools/microbench/microbench

# 1. parity on hardware, file by file (no -x: every file reports; the blind files are the zz*/zzz* ones)
for f in tests/test_gpu_forcing.py tests/test_gpu_marking.py tests/test_gpu_stress.py tests/test_gpu_motion.py \
         tests/test_gpu_multirank.py tests/test_gpu_zz_device_api.py tests/test_gpu_zz_reference_vectors.py tests/test_host_mirror.py \
         tests/test_zz_gpu_shim_dropin.py tests/test_zz_gpu_validation_cases.py tests/test_zzz_gpu_active_loop5.py \
         tests/test_zzz_gpu_mark_variant.py tests/test_zzz_gpu_soft_sphere.py tests/test_zzz_gpu_v5_kernels.py; do
  [ -f "$f" ] && leg "pytest_$(basename "$f" .py)" 900 $PY -m pytest "$f" -m gpu -q -p no:cacheprovider
done
# the slab-decomposed path with every blind variant forced through the developer switches
leg pytest_multirank_blind_variants 900 env IBM_KERNEL_VARIANT=6,5 IBM_MARK_VARIANT=1 IBM_ACTIVE_LOOP5=1 IBM_FUSED_ACTIVE_LOOP5=1 \
    $PY -m pytest tests/test_gpu_multirank.py -m gpu -q -p no:cacheprovider
leg smoke 300 $PY -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"

# 2. the bench line the driver will take (probe decides), then each generation forced, then the loop-5 variants
leg bench_probe 900 $PY bench.py --steps 5 --warmup 3
leg bench_defaults 600 $PY bench.py --steps 5 --warmup 3 --no-probe --no-cpu-baseline
leg bench_v5 600 $PY bench.py --steps 5 --warmup 3 --variant 5 --no-cpu-baseline --no-e2e
leg bench_v6 600 $PY bench.py --steps 5 --warmup 3 --variant 6 --no-cpu-baseline --no-e2e
leg bench_v5_split 600 $PY bench.py --steps 5 --warmup 3 --variant 5 --split 1 --no-cpu-baseline --no-e2e
leg bench_v2_split 600 $PY bench.py --steps 5 --warmup 3 --no-probe --split 1 --no-cpu-baseline --no-e2e
leg bench_v2_split_fused 600 $PY bench.py --steps 5 --warmup 3 --no-probe --split 2 --no-cpu-baseline --no-e2e
leg bench_v5_split_fused 600 $PY bench.py --steps 5 --warmup 3 --variant 5 --split 2 --no-cpu-baseline --no-e2e
leg bench_v2_mark2 600 $PY bench.py --steps 5 --warmup 3 --no-probe --mark 1 --no-cpu-baseline --no-e2e
leg bench_reference 600 $PY bench.py --impl reference --steps 2 --warmup 1

# 3. launch list of a step (v5 + split) and one full capture of each fifth-generation kernel (one GPU, after the runs above exited)
if command -v ncu > /dev/null; then
  leg ncu_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
      --log-file "$OUT/launches_v5_split.csv" $PY bench.py --steps 1 --warmup 3 --variant 5 --split 1 --no-cpu-baseline --no-e2e
  leg ncu_forcing_v5 900 ncu --set full --clock-control none --import-source on -k regex:k_forcing_v5 -s 3 -c 1 \
      -o "$OUT/forcing_v5" -f $PY bench.py --steps 1 --warmup 3 --variant 5 --no-cpu-baseline --no-e2e
  leg ncu_forcing_v6 900 ncu --set full --clock-control none --import-source on -k regex:k_forcing_v5 -s 3 -c 1 \
      -o "$OUT/forcing_v6" -f $PY bench.py --steps 1 --warmup 3 --variant 6 --no-cpu-baseline --no-e2e
  leg ncu_stress_v5 900 ncu --set full --clock-control none --import-source on -k regex:k_stress_v5 -s 3 -c 1 \
      -o "$OUT/stress_v5" -f $PY bench.py --steps 1 --warmup 3 --variant 5 --no-cpu-baseline --no-e2e
fi
grep -h '^{' "$OUT"/bench_*.log > "$OUT/bench_lines.jsonl" 2>/dev/null
python tools/summarize_first_call.py "$OUT" > "$OUT/SUMMARY.md" 2>&1
cat "$OUT/summary.txt"
