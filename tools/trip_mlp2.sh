# one-GPU trip: two-tile fused MLP kernel, spinning vs suspending mbarrier waits on the fc1 -> GELU -> fc2 chain
mkdir -p gpurun_out
for sp in 0 1 2 3; do
  RVK_MLP2_SPIN=$sp timeout 120 python tools/kbench_mlp.py 4 1024 20 proj 2>&1 | head -1
done
for sp in 1 3; do
  RVK_MLP2_SPIN=$sp timeout 300 python -m pytest tests/test_gpu_mlp_fused.py -x -q -m gpu -k "attn_proj and 4" 2>&1 | tail -1
done
for sp in 0 1 2 3; do
  RVK_MLP2_SPIN=$sp timeout 120 python tools/kbench_mlp.py 4 1024 20 proj 2>&1 | head -1
done
