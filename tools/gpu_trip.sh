#!/bin/bash
# Runs on the GPU box under gpurun: every GPU test file in its own process (a trapped kernel must not take
# the other files down), then smoke and a short bench.  Logs land in gpurun_out/.
# usage: bash tools/gpu_trip.sh [tests|notests] [bench|nobench] [ncu|noncu]
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.used --format=csv > gpurun_out/nvsmi.txt 2>&1
if [ "${1:-tests}" = "tests" ]; then
for f in ${GPU_TEST_FILES:-test_gpu_mlp_fused test_gpu_gemm test_gpu_encoder_kernels test_gpu_heads test_gpu_model}; do
  echo "=== $f" | tee -a gpurun_out/summary.txt
  timeout 420 python -m pytest tests/$f.py -q -m gpu --no-header -rN -s > gpurun_out/$f.log 2>&1
  echo "exit $?" | tee -a gpurun_out/summary.txt
  tail -n 3 gpurun_out/$f.log | tee -a gpurun_out/summary.txt
done
echo "=== smoke" | tee -a gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary.txt
tail -n 2 gpurun_out/smoke.log | tee -a gpurun_out/summary.txt
fi
if [ "${2:-bench}" = "bench" ]; then
  echo "=== bench" | tee -a gpurun_out/summary.txt
  timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; rc=$?; echo "exit $rc" | tee -a gpurun_out/summary.txt
  tail -n 3 gpurun_out/bench.log | tee -a gpurun_out/summary.txt
  if [ "${3:-noncu}" = "ncu" ] && [ $rc -eq 0 ]; then
    echo "=== ncu launch list" | tee -a gpurun_out/summary.txt
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${NCU_SKIP:-290} -c ${NCU_COUNT:-100} --csv \
      --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
    echo "exit $?" | tee -a gpurun_out/summary.txt
    tail -n 2 gpurun_out/ncu.log | tee -a gpurun_out/summary.txt
  fi
fi
