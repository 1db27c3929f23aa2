#!/bin/bash
# Runs on the GPU box under gpurun: every GPU test file in its own process (a trapped kernel must not take
# the other files down), then smoke and a short bench.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.used --format=csv > gpurun_out/nvsmi.txt 2>&1
for f in ${TEST_FILES:-test_gpu_gemm test_gpu_encoder_kernels test_gpu_heads test_gpu_model}; do
  echo "=== $f" | tee -a gpurun_out/summary.txt
  timeout ${TEST_TIMEOUT:-420} python -m pytest tests/$f.py -q -m gpu -x --no-header -rN ${PYTEST_ARGS:-} > gpurun_out/$f.log 2>&1
  echo "exit $?" | tee -a gpurun_out/summary.txt
  tail -n 3 gpurun_out/$f.log | tee -a gpurun_out/summary.txt
done
if [ -z "$SKIP_BENCH" ]; then
  echo "=== smoke" | tee -a gpurun_out/summary.txt
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary.txt
  tail -n 2 gpurun_out/smoke.log | tee -a gpurun_out/summary.txt
  echo "=== bench" | tee -a gpurun_out/summary.txt
  timeout 600 python bench.py --steps ${BENCH_STEPS:-10} --warmup 3 > gpurun_out/bench.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary.txt
  tail -n 3 gpurun_out/bench.log | tee -a gpurun_out/summary.txt
fi
