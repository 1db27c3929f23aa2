GPU_TEST_FILES="test_gpu_encoder_kernels test_gpu_heads test_gpu_model" bash tools/gpu_trip_r2.sh tests nobench noncu
python bench.py --mode train --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_train.log 2>&1
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_train.log') if l.startswith('{')][-1])
print('train', d['value'], d['ms_per_step'], d['phases_ms'], d['gpu_launches'], 'cutmix', d['cutmix']['value'], 'frozen', d['frozen_backbone']['value'])
print({k:(round(v['us_per_launch'],1), v['launches_per_step'], round(v['frac'] or 0,3)) for k,v in d['kernels'].items()})
PY
