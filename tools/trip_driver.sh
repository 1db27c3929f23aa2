# what the driver runs at round end, in the same form: the whole GPU suite in ONE process, smoke(), both bench arms
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/driver_pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/driver_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
if [ "${1:-bench}" = "bench" ]; then
  t0=$SECONDS; timeout 900 python bench.py --impl reference > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref exit $? in $((SECONDS - t0)) s"; tail -c 400 gpurun_out/bench_ref.log
  t0=$SECONDS; timeout 900 python bench.py > gpurun_out/bench_all.log 2> gpurun_out/bench_all.err; echo "bench exit $? in $((SECONDS - t0)) s"; tail -c 300 gpurun_out/bench_all.log
fi
