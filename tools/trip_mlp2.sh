# one-GPU trip: parity of the two-tiles-in-flight fused MLP kernel (cta_group 4) + microbench / clock trace against cta_group 2
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_mlp_fused.py -x -q -m gpu -k "attn_proj" > gpurun_out/mlp2_tests.log 2>&1
echo "tests exit $?"; tail -4 gpurun_out/mlp2_tests.log
timeout 120 python tools/kbench_mlp.py 2 1024 20 proj 2>&1 | head -1
for q in ${MLP2_PROJQS:-1 2 3}; do
  RVK_MLP2_PROJQ=$q timeout 120 python tools/kbench_mlp.py 4 1024 20 proj 2>&1 | head -1
done
MLP_TRACE=1 timeout 120 python tools/kbench_mlp.py 4 1024 20 proj > gpurun_out/mlp_trace_g4.log 2>&1; echo "trace exit $?"
