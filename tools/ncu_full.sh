#!/bin/bash
# one --set full capture of the GEMM family (5 launches = all epilogue modes of one block) and of attention
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_nt -s 245 -c 5 -o gpurun_out/prof_gemm -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1
echo "gemm exit $?"
ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 37 -c 1 -o gpurun_out/prof_attn -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_attn.log 2>&1
echo "attn exit $?"
ls -la gpurun_out/*.ncu-rep
