mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_all_n2.log 2>&1; echo "exit $?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_all_n2.log') if l.startswith('{')][-1])
t=d['train']; print('N=2 infer', round(d['value']), 'train', round(t['value']), t['phases_ms'], t['allreduce'])
PY
tail -3 gpurun_out/bench_all_n2.log | cut -c1-300
