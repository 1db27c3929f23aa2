GPU_TEST_FILES="test_gpu_heads test_gpu_optim test_gpu_api_misc test_gpu_parity_full test_gpu_model" bash tools/gpu_trip_r2.sh tests nobench noncu
python bench.py --mode kan --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_kan.log 2>&1; tail -c 400 gpurun_out/bench_kan.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_kan.csv python bench.py --mode kan --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_kan.log 2>&1
python bench.py --mode train --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train.log 2>&1; tail -c 300 gpurun_out/bench_train.log
