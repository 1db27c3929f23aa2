GPU_TEST_FILES="test_gpu_heads test_gpu_parity_full test_gpu_model test_gpu_api_misc test_gpu_optim" bash tools/gpu_trip_r2.sh tests nobench noncu
python bench.py --mode kan --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_kan.log 2>&1
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_kan.log') if l.startswith('{')][-1])
print('kan', d['detail'], {k:round(v['us_per_launch'],1) for k,v in d['kernels'].items()})
PY
for ft in 1 0; do
RVK_FUSED_TAIL=$ft python bench.py --mode train --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_train_ft$ft.log 2>&1
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_train_ft$ft.log') if l.startswith('{')][-1])
print('train fused_tail=$ft', d['value'], d['ms_per_step'], d['phases_ms'], d['gpu_launches'], 'cutmix', d['cutmix']['value'], 'frozen', d['frozen_backbone']['value'])
print({k:(round(v['us_per_launch'],1), v['launches_per_step']) for k,v in d['kernels'].items()})
PY
done
