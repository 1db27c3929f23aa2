# one-GPU trip: parity of the two-tiles-in-flight fused MLP kernel (cta_group 4) + microbench / clock trace against cta_group 2
mkdir -p gpurun_out
for v in ${MLP2_VARS:-0 1 2 3}; do
  RVK_MLP2_VAR=$v timeout 300 python -m pytest tests/test_gpu_mlp_fused.py -x -q -m gpu -k "attn_proj and 4" > gpurun_out/mlp2_tests_v$v.log 2>&1
  echo "tests var=$v exit $?"; tail -1 gpurun_out/mlp2_tests_v$v.log
done
timeout 120 python tools/kbench_mlp.py 2 1024 20 proj 2>&1 | head -1
for v in ${MLP2_VARS:-0 1 2 3}; do
  for c in ${MLP2_PROJQS:-4}; do
    RVK_MLP2_VAR=$v RVK_MLP2_PROJQ=$c MLP_TRACE=1 timeout 120 python tools/kbench_mlp.py 4 1024 20 proj > gpurun_out/mlp_trace_g4_v${v}_q$c.log 2>&1
    echo "var=$v projq=$c exit $?"; head -1 gpurun_out/mlp_trace_g4_v${v}_q$c.log
  done
done
