# one-GPU trip: host topology probe, KAN bench (eager + graph replay), compute-sanitizer passes over the small-shape tests
mkdir -p gpurun_out
{
  echo "== nvidia-smi topo -m"; nvidia-smi topo -m
  echo "== lscpu"; lscpu | head -30
  echo "== numa nodes"; ls /sys/devices/system/node/ 2>/dev/null; for n in /sys/devices/system/node/node*; do echo $n $(cat $n/cpulist); done
  echo "== gpu pci numa"; for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ]; then echo $d $(cat $d/class) numa=$(cat $d/numa_node) local_cpus=$(cat $d/local_cpulist); fi; done
  echo "== affinity"; python -c "import os; print(len(os.sched_getaffinity(0)), os.cpu_count())"
  echo "== meminfo"; head -3 /proc/meminfo
} > gpurun_out/topo.txt 2>&1
python bench.py --mode kan --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_kan.log 2>&1
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_kan.log') if l.startswith('{')][-1])
print('kan', d['ms_per_step'], d['ms_per_step_eager_launch'], d['detail'], d['roofline']['frac'])
PY
for t in test_gpu_heads test_gpu_optim; do
  timeout 500 compute-sanitizer --tool memcheck --target-processes all --error-exitcode 7 python -m pytest tests/$t.py -x -q -m gpu > gpurun_out/memcheck_$t.log 2>&1
  echo "memcheck $t exit $?"; tail -4 gpurun_out/memcheck_$t.log
done
timeout 400 compute-sanitizer --tool racecheck --target-processes all --error-exitcode 7 python -m pytest tests/test_gpu_heads.py -x -q -m gpu -k "fused" > gpurun_out/racecheck_heads.log 2>&1
echo "racecheck heads exit $?"; tail -4 gpurun_out/racecheck_heads.log
