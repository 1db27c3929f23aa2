# experiment: two-stage / eight-slot gemm_nt for every K = 192 GEMM vs the default (DGELU only)
mkdir -p gpurun_out
for v in 0 2; do
  RVK_NT_STAGES=$v python bench.py --mode train --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_train_st$v.log 2>&1
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_train_st$v.log') if l.startswith('{')][-1])
print('stages env $v: train', round(d['value']), d['ms_per_step'], 'gemm_nt', round(d['kernels']['gemm_nt_kernel']['ms_per_step'],3))
PY
  RVK_NT_STAGES=$v python bench.py --mode infer --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_infer_st$v.log 2>&1
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_infer_st$v.log') if l.startswith('{')][-1])
print('stages env $v: infer', round(d['value']), d['ms_per_step'], 'gemm_nt', round(d['roofline']['kernels']['gemm_nt_kernel']['ms_per_step'],3))
PY
done
RVK_NT_STAGES=2 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train_st2.csv python bench.py --mode train --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_train_st2.log 2>&1
python tools/launch_summary.py gpurun_out/launches_train_st2.csv | head -12
