# ncu --set full of the training-only GEMM epilogue modes (one steady-state launch each) -> gpurun_out/prof_gemm_nt_mode*.ncu-rep
mkdir -p gpurun_out
BT="python bench.py --mode train --steps 1 --warmup 3 --no-cpu-baseline"
$BT > gpurun_out/plain_train.log 2>&1 || { echo "plain run failed"; tail gpurun_out/plain_train.log; exit 1; }
NCU="ncu --set full --clock-control none --import-source on -f --kernel-name-base demangled"
for m in ${MODES:-1 2}; do
  timeout 600 $NCU -k "regex:gemm_nt_kernel<.int.192, .int.$m," -s 14 -c 1 -o gpurun_out/prof_gemm_nt_mode$m $BT > gpurun_out/ncu_gemm_nt_mode$m.log 2>&1; echo "mode $m exit $?"
done
ls -la gpurun_out/prof_gemm_nt_mode*.ncu-rep
