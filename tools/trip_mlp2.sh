# one-GPU trip: parity of the two-tiles-in-flight fused MLP kernel (cta_group 4) + microbench / clock trace against cta_group 2
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_mlp_fused.py -x -q -m gpu -k "attn_proj" > gpurun_out/mlp2_tests.log 2>&1
echo "tests exit $?"; tail -5 gpurun_out/mlp2_tests.log
timeout 120 python tools/kbench_mlp.py 2 1024 20 proj 2>&1 | head -1
for c in ${MLP2_PROJCS:-0 1 2 3}; do
  RVK_MLP2_PROJC=$c MLP_TRACE=1 timeout 120 python tools/kbench_mlp.py 4 1024 20 proj > gpurun_out/mlp_trace_g4_c$c.log 2>&1
  echo "projc=$c exit $?"; head -1 gpurun_out/mlp_trace_g4_c$c.log
done
