#!/bin/bash
# Round-2 GPU trip: every GPU test file in its own process, smoke, the default bench line, then (optionally) ncu launch
# lists and --set full captures of the training / KAN kernels.  Logs land in gpurun_out/.
# usage: bash tools/gpu_trip_r2.sh [tests|notests] [bench|nobench] [ncu|noncu]
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.used --format=csv > gpurun_out/nvsmi.txt 2>&1
if [ "${1:-tests}" = "tests" ]; then
for f in ${GPU_TEST_FILES:-test_gpu_mlp_fused test_gpu_gemm test_gpu_encoder_kernels test_gpu_heads test_gpu_model test_gpu_api_misc test_gpu_parity_full test_gpu_dropin_flow}; do
  echo "=== $f" | tee -a gpurun_out/summary.txt
  timeout 600 python -m pytest tests/$f.py -q -m gpu --no-header -rN -s > gpurun_out/$f.log 2>&1
  echo "exit $?" | tee -a gpurun_out/summary.txt
  tail -n 3 gpurun_out/$f.log | tee -a gpurun_out/summary.txt
done
echo "=== smoke" | tee -a gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary.txt
tail -n 2 gpurun_out/smoke.log | tee -a gpurun_out/summary.txt
fi
if [ "${2:-bench}" = "bench" ]; then
  echo "=== bench (default line)" | tee -a gpurun_out/summary.txt
  timeout 900 python bench.py --steps ${BENCH_STEPS:-20} --warmup 5 > gpurun_out/bench_all.log 2>&1; rc=$?; echo "exit $rc" | tee -a gpurun_out/summary.txt
  tail -c 600 gpurun_out/bench_all.log | tee -a gpurun_out/summary.txt
fi
if [ "${3:-noncu}" = "ncu" ]; then
  NCU_LIST="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
  echo "=== ncu launch lists" | tee -a gpurun_out/summary.txt
  BT="python bench.py --mode train --steps 1 --warmup 3 --no-cpu-baseline"
  $BT > gpurun_out/plain_train.log 2>&1 && timeout 900 $NCU_LIST --log-file gpurun_out/launches_train.csv $BT > gpurun_out/ncu_train.log 2>&1
  echo "train exit $?" | tee -a gpurun_out/summary.txt
  BK="python bench.py --mode kan --steps 2 --warmup 3 --no-cpu-baseline"
  $BK > gpurun_out/plain_kan.log 2>&1 && timeout 600 $NCU_LIST --log-file gpurun_out/launches_kan.csv $BK > gpurun_out/ncu_kan.log 2>&1
  echo "kan exit $?" | tee -a gpurun_out/summary.txt
  BI="python bench.py --mode infer --steps 2 --warmup 3 --no-cpu-baseline"
  $BI > gpurun_out/plain_infer.log 2>&1 && timeout 600 $NCU_LIST -s 250 -c 100 --log-file gpurun_out/launches_infer.csv $BI > gpurun_out/ncu_infer.log 2>&1
  echo "infer exit $?" | tee -a gpurun_out/summary.txt
  NCU="ncu --set full --clock-control none --import-source on -f"
  for k in ${NCU_KAN_KERNELS:-kan_fwd_tc kan_bwd_x_tc kan_bwd_w_tc}; do
    timeout 600 $NCU -k regex:$k -s 4 -c 1 -o gpurun_out/prof_$k $BK > gpurun_out/ncu_$k.log 2>&1; echo "$k exit $?" | tee -a gpurun_out/summary.txt
  done
  for k in ${NCU_TRAIN_KERNELS:-gemm_tn layernorm_bwd attn_bwd_tc}; do
    timeout 600 $NCU -k regex:$k -s 60 -c 1 -o gpurun_out/prof_$k $BT > gpurun_out/ncu_$k.log 2>&1; echo "$k exit $?" | tee -a gpurun_out/summary.txt
  done
  timeout 600 $NCU -k regex:heads_fused_kernel -s 3 -c 1 -o gpurun_out/prof_heads_fused $BI > gpurun_out/ncu_heads_fused.log 2>&1; echo "heads_fused exit $?" | tee -a gpurun_out/summary.txt
  ls -la gpurun_out/*.ncu-rep | tee -a gpurun_out/summary.txt
fi
