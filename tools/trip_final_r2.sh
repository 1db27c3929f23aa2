# round-2 final trip: driver-style run (all GPU tests in one process, smoke, both bench arms), the one-tile kernel for comparison on
# the same box, the inference launch list and the --set full capture of the two-tile fused MLP kernel
bash tools/trip_driver.sh bench
RVK_MLP_CTA_GROUP=2 timeout 300 python bench.py --mode infer --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_infer_g2.log 2> gpurun_out/bench_infer_g2.err; echo "infer one-tile kernel exit $?"
timeout 300 python bench.py --mode infer --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_infer_g4.log 2> gpurun_out/bench_infer_g4.err; echo "infer two-tile kernel exit $?"
NCU_LIST="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
B="python bench.py --mode infer --steps 2 --warmup 3 --no-cpu-baseline"
timeout 600 $NCU_LIST -s 250 -c 100 --log-file gpurun_out/launches_infer.csv $B > gpurun_out/ncu_infer.log 2>&1; echo "list infer $?"
python tools/launch_summary.py gpurun_out/launches_infer.csv > gpurun_out/launches_infer_summary.txt 2>&1; head -6 gpurun_out/launches_infer_summary.txt
NCU="ncu --set full --clock-control none --import-source on -f --kernel-name-base demangled"
timeout 600 $NCU -k "regex:mlp_fused2" -s 40 -c 1 -o gpurun_out/prof_mlp2 python bench.py --mode infer --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_mlp2.log 2>&1; echo "mlp2 ncu exit $?"
