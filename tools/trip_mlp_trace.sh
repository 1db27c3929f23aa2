# one-GPU trip: fused projection + MLP kernel alone at the inference batch, with the clock64 event log of CTA 0
mkdir -p gpurun_out
MLP_TRACE=1 timeout 120 python tools/kbench_mlp.py 2 1024 20 proj > gpurun_out/mlp_trace.log 2>&1
echo "exit $?"; cut -c1-400 gpurun_out/mlp_trace.log | head -5
