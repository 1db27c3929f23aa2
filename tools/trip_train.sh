# one-GPU trip for training-path changes: the tests that cover the trunk, the train bench, the launch list of one step
GPU_TEST_FILES="${GPU_TEST_FILES:-test_gpu_gemm test_gpu_encoder_kernels test_gpu_model test_gpu_parity_full}" bash tools/gpu_trip_r2.sh tests nobench noncu
python bench.py --mode train --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_train.log 2>&1
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_train.log') if l.startswith('{')][-1])
print('train', round(d['value']), d['ms_per_step'], {k:(round(v['us_per_launch'],1), round(v['ms_per_step'],3)) for k,v in d['kernels'].items()})
PY
BT="python bench.py --mode train --steps 1 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv $BT > gpurun_out/ncu_train.log 2>&1
python tools/launch_summary.py gpurun_out/launches_train.csv > gpurun_out/launches_train_summary.txt; head -14 gpurun_out/launches_train_summary.txt
