#!/bin/bash
# --set full captures of the hot inference kernels (one launch each, steady state) -> gpurun_out/prof_*.ncu-rep
# Run under gpurun; summarise here with tools/ncu_summary.py and tools/ncu_hot.py, commit the text under profiles/.
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail gpurun_out/plain.log; exit 1; }
NCU="ncu --set full --clock-control none --import-source on -f"
# launch order per forward: im2col, patch GEMM, 12 x (qkv gemm<..,0,..>, attention, proj gemm<..,4,..>, mlp_fused), ...; skip 3 warm-up forwards
$NCU -k regex:mlp_fused -s 40 -c 1 -o gpurun_out/prof_mlp $B > gpurun_out/ncu_mlp.log 2>&1; echo "mlp exit $?"
$NCU -k regex:attn_fwd -s 40 -c 1 -o gpurun_out/prof_attn $B > gpurun_out/ncu_attn.log 2>&1; echo "attn exit $?"
$NCU -k regex:gemm_nt -s 78 -c 13 -o gpurun_out/prof_gemm $B > gpurun_out/ncu_gemm.log 2>&1; echo "gemm exit $?"
ls -la gpurun_out/*.ncu-rep
