mkdir -p gpurun_out
RVK_TN_SIDE_STREAM=1 timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_parity_full.py -q -m gpu --no-header -rN -k "vjp or gradients or train_step or torchvision" > gpurun_out/test_side_stream.log 2>&1; echo "side-stream tests exit $?"; tail -3 gpurun_out/test_side_stream.log
for ss in 0 1 0 1; do
RVK_TN_SIDE_STREAM=$ss python bench.py --mode train --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_train_ss$ss.log 2>&1
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_train_ss$ss.log') if l.startswith('{')][-1])
print('side_stream=$ss train', round(d['value']), round(d['ms_per_step'],3), d['phases_ms'], 'cutmix', round(d['cutmix']['value']))
PY
done
