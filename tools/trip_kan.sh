GPU_TEST_FILES="test_gpu_heads test_gpu_parity_full" bash tools/gpu_trip_r2.sh tests nobench noncu
python bench.py --mode kan --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_kan.log 2>&1
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_kan.log') if l.startswith('{')][-1])
print('kan', d['detail'], {k:round(v['us_per_launch'],1) for k,v in d['kernels'].items()}, d['roofline']['frac'])
PY
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_kan.csv python bench.py --mode kan --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_kan.log 2>&1
python tools/launch_summary.py gpurun_out/launches_kan.csv > gpurun_out/launches_kan_summary.txt; head -8 gpurun_out/launches_kan_summary.txt
