# round-2 final trip: driver-style run (all GPU tests in one process, smoke, both bench arms) + the inference launch list
bash tools/trip_driver.sh bench
NCU_LIST="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 600 $NCU_LIST -s 250 -c 100 --log-file gpurun_out/launches_infer.csv python bench.py --mode infer --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_infer.log 2>&1; echo "list infer $?"
python tools/launch_summary.py gpurun_out/launches_infer.csv > gpurun_out/launches_infer_summary.txt 2>&1; head -12 gpurun_out/launches_infer_summary.txt
