# ncu --set full of the two-tile fused MLP kernel (one steady-state launch of the inference forward)
mkdir -p gpurun_out
B="python bench.py --mode infer --steps 1 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain_infer.log 2>&1 || { echo "plain run failed"; tail gpurun_out/plain_infer.log; exit 1; }
NCU="ncu --set full --clock-control none --import-source on -f --kernel-name-base demangled"
timeout 600 $NCU -k "regex:mlp_fused2" -s 40 -c 1 -o gpurun_out/prof_mlp2 $B > gpurun_out/ncu_mlp2.log 2>&1; echo "mlp2 exit $?"
ls -la gpurun_out/prof_mlp2.ncu-rep
