# ncu --set full of the three tensor-core KAN kernels (one steady-state launch each) -> gpurun_out/prof_kan_*.ncu-rep
mkdir -p gpurun_out
BK="python bench.py --mode kan --steps 2 --warmup 3 --no-cpu-baseline"
$BK > gpurun_out/plain_kan.log 2>&1 || { echo "plain run failed"; tail gpurun_out/plain_kan.log; exit 1; }
NCU="ncu --set full --clock-control none --import-source on -f"
for k in ${NCU_KAN_KERNELS:-kan_fwd_tc kan_bwd_x_tc kan_bwd_w_tc}; do
  timeout 600 $NCU -k regex:$k -s 4 -c 1 -o gpurun_out/prof_$k $BK > gpurun_out/ncu_$k.log 2>&1; echo "$k exit $?"
done
ls -la gpurun_out/prof_kan*.ncu-rep
