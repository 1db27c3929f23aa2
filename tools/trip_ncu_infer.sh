# ncu --set full of the inference qkv GEMM and the attention forward (one steady-state launch each)
mkdir -p gpurun_out
B="python bench.py --mode infer --steps 1 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain_infer.log 2>&1 || { echo "plain run failed"; tail gpurun_out/plain_infer.log; exit 1; }
NCU="ncu --set full --clock-control none --import-source on -f --kernel-name-base demangled"
timeout 600 $NCU -k "regex:gemm_nt_kernel<.int.192, .int.0," -s 40 -c 1 -o gpurun_out/prof_gemm_qkv_infer $B > gpurun_out/ncu_gemm_qkv_infer.log 2>&1; echo "qkv exit $?"
timeout 600 $NCU -k "regex:attn_fwd_tc" -s 40 -c 1 -o gpurun_out/prof_attn_infer $B > gpurun_out/ncu_attn_infer.log 2>&1; echo "attn exit $?"
ls -la gpurun_out/prof_gemm_qkv_infer.ncu-rep gpurun_out/prof_attn_infer.ncu-rep
