# one-GPU trip: inference forward (batch 1024) with the one-tile (2) and the two-tile (4) fused MLP kernel, same box; model-level parity tests with 4
mkdir -p gpurun_out
for g in 2 4 2 4; do
  RVK_MLP_CTA_GROUP=$g timeout 300 python bench.py --mode infer --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_infer_g$g.log 2> gpurun_out/bench_infer_g$g.err
  echo "infer g=$g exit $?"
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_infer_g$g.log') if l.startswith('{')][-1])
print('g=$g', d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('e2e',{}).get('value'))
PY
done
RVK_MLP_CTA_GROUP=4 timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_parity_full.py -x -q -m gpu > gpurun_out/tests_model_g4.log 2>&1
echo "model tests g=4 exit $?"; tail -3 gpurun_out/tests_model_g4.log
