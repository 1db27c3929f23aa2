#!/bin/bash
# N-GPU trip (gpurun --gpus N): the default bench line through torchrun exactly as the driver launches it, then the train
# reference arm launched the same way (rank 0 works, the others exit 0), then the multi-GPU tests.
N=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $RUN bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_all_n$N.log 2>&1; echo "all exit $?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_all_n$N.log') if l.startswith('{')][-1])
print('N=$N infer', round(d['value']), 'e2e', round(d['e2e']['value']), 'u8', round(d['e2e_uint8_input']['value']))
t=d['train']; print('train', round(t['value']), t['ms_per_step'], t['phases_ms'], t['allreduce'])
print('kan', round(d['kan']['value']), 'sweep best', round(d['sweep']['value']))
PY
timeout 600 $RUN bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.log 2>&1; echo "reference arm exit $?"; tail -c 300 gpurun_out/bench_ref_n$N.log
timeout 600 python -m pytest tests/test_gpu_dist.py -q -m gpu -s > gpurun_out/test_gpu_dist.log 2>&1; echo "dist test exit $?"; tail -3 gpurun_out/test_gpu_dist.log
timeout 300 python -m pytest tests/test_gpu_optim.py -q -m gpu -k fast_trainer > gpurun_out/test_fast_trainer.log 2>&1; echo "fast trainer test exit $?"; tail -2 gpurun_out/test_fast_trainer.log
