GPU_TEST_FILES="test_gpu_heads test_gpu_optim test_gpu_parity_full test_gpu_model" bash tools/gpu_trip_r2.sh tests nobench noncu
for n in 2 4 8; do
  RVK_KAN_SMALL_CTAS_PER_SM=$n python bench.py --mode kan --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_kan_$n.log 2>&1
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_kan_$n.log') if l.startswith('{')][-1])
print('ctas/sm $n', d['detail'], {k:round(v['us_per_launch'],1) for k,v in d['kernels'].items()})
PY
done
ncu --set full --clock-control none --import-source on -f -k regex:kan_small_bwd -s 3 -c 2 -o gpurun_out/prof_kan_small_bwd python bench.py --mode kan --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_kan_small_bwd.log 2>&1; echo "ncu exit $?"
