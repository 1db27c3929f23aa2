GPU_TEST_FILES="test_gpu_heads test_gpu_api_misc test_gpu_model" bash tools/gpu_trip_r2.sh tests nobench noncu
python bench.py --mode train --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_train.log 2>&1
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_train.log') if l.startswith('{')][-1])
print('train', round(d['value']), round(d['ms_per_step'],3), d['phases_ms'], 'roofline', {k:v for k,v in d['roofline'].items() if k in ('achieved','frac','traffic','us_per_launch','share_of_step')})
print({k:(round(v['us_per_launch'],1), v['launches_per_step']) for k,v in d['kernels'].items()})
PY
