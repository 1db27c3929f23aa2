# final validation + evidence: all tests, default bench line, launch lists, --set full captures; the reports are summarised ON the
# box (gpurun merges at most 64 MiB back): text summaries land in gpurun_out/profiles_r2/, only three reports travel.
# usage: bash tools/gpu_profiles_r2.sh [tests|notests]
GPU_TEST_FILES="test_gpu_mlp_fused test_gpu_gemm test_gpu_encoder_kernels test_gpu_heads test_gpu_model test_gpu_api_misc test_gpu_parity_full test_gpu_dropin_flow test_gpu_optim" bash tools/gpu_trip_r2.sh ${1:-tests} bench noncu
NCU_LIST="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
BT="python bench.py --mode train --steps 1 --warmup 3 --no-cpu-baseline"
BK="python bench.py --mode kan --steps 2 --warmup 3 --no-cpu-baseline"
BI="python bench.py --mode infer --steps 2 --warmup 3 --no-cpu-baseline"
timeout 900 $NCU_LIST --log-file gpurun_out/launches_train.csv $BT > gpurun_out/ncu_train.log 2>&1; echo "list train $?"
timeout 600 $NCU_LIST --log-file gpurun_out/launches_kan.csv $BK > gpurun_out/ncu_kan.log 2>&1; echo "list kan $?"
timeout 600 $NCU_LIST -s 250 -c 100 --log-file gpurun_out/launches_infer.csv $BI > gpurun_out/ncu_infer.log 2>&1; echo "list infer $?"
NCU="ncu --set full --clock-control none --import-source on -f"
# KAN: the first matching launches after warm-up belong to the 192 -> 64 layer of the [192,64,1] stack
timeout 600 $NCU -k regex:kan_fwd_tc -s 3 -c 1 -o gpurun_out/prof_kan_fwd_tc $BK > gpurun_out/ncu_a.log 2>&1; echo "kan_fwd_tc $?"
timeout 600 $NCU -k regex:kan_bwd_x_tc -s 3 -c 1 -o gpurun_out/prof_kan_bwd_x_tc $BK > gpurun_out/ncu_b.log 2>&1; echo "kan_bwd_x_tc $?"
timeout 600 $NCU -k regex:kan_bwd_w_tc -s 3 -c 1 -o gpurun_out/prof_kan_bwd_w_tc $BK > gpurun_out/ncu_c.log 2>&1; echo "kan_bwd_w_tc $?"
timeout 600 $NCU -k regex:kan_small_bwd -s 3 -c 1 -o gpurun_out/prof_kan_small_bwd $BK > gpurun_out/ncu_d.log 2>&1; echo "kan_small_bwd $?"
timeout 600 $NCU -k regex:kan_small_fwd -s 3 -c 1 -o gpurun_out/prof_kan_small_fwd $BK > gpurun_out/ncu_e.log 2>&1; echo "kan_small_fwd $?"
for k in gemm_tn layernorm_bwd attn_bwd_tc; do
  timeout 600 $NCU -k regex:$k -s 60 -c 1 -o gpurun_out/prof_$k $BT > gpurun_out/ncu_$k.log 2>&1; echo "$k $?"
done
NCUD="$NCU --kernel-name-base demangled"
timeout 600 $NCUD -k "regex:gemm_nt_kernel<.int.192, .int.1," -s 14 -c 1 -o gpurun_out/prof_gemm_nt_gelu $BT > gpurun_out/ncu_k.log 2>&1; echo "gemm_nt gelu $?"
timeout 600 $NCUD -k "regex:gemm_nt_kernel<.int.192, .int.2," -s 14 -c 1 -o gpurun_out/prof_gemm_nt_dgelu $BT > gpurun_out/ncu_l.log 2>&1; echo "gemm_nt dgelu $?"
timeout 600 $NCU -k regex:heads_train_bwd -s 3 -c 1 -o gpurun_out/prof_heads_train_bwd $BT > gpurun_out/ncu_f.log 2>&1; echo "heads_train_bwd $?"
timeout 600 $NCU -k regex:heads_fused_kernel -s 3 -c 1 -o gpurun_out/prof_heads_train_fwd $BT > gpurun_out/ncu_g.log 2>&1; echo "heads_train_fwd $?"
timeout 600 $NCU -k regex:optim_update -s 3 -c 1 -o gpurun_out/prof_optim_update $BT > gpurun_out/ncu_h.log 2>&1; echo "optim_update $?"
timeout 600 $NCU -k regex:heads_fused_kernel -s 3 -c 1 -o gpurun_out/prof_heads_fused $BI > gpurun_out/ncu_i.log 2>&1; echo "heads_fused $?"
timeout 600 $NCU -k regex:mlp_fused -s 40 -c 1 -o gpurun_out/prof_mlp $BI > gpurun_out/ncu_j.log 2>&1; echo "mlp $?"
ls -la gpurun_out/*.ncu-rep
python tools/make_profiles.py r2 > gpurun_out/make_profiles.log 2>&1; echo "make_profiles $?"
mkdir -p gpurun_out/profiles_r2 gpurun_out/keep
cp profiles/r2_* profiles/roofline_traffic.json gpurun_out/profiles_r2/
for k in kan_fwd_tc gemm_nt_dgelu mlp; do mv gpurun_out/prof_$k.ncu-rep gpurun_out/keep/ 2>/dev/null; done
rm -f gpurun_out/prof_*.ncu-rep gpurun_out/launches_*.csv
mv gpurun_out/keep/*.ncu-rep gpurun_out/ 2>/dev/null; rmdir gpurun_out/keep
du -sh gpurun_out
