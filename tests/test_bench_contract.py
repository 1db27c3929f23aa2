"""bench.py contract checks that run without a GPU: the reference arm (the reference's CPU algorithm on the host cores) prints one
JSON line with the same metric / unit / config keys as our arm, plus the cpu_baseline and e2e objects the tier asks for."""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra):
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1', '--no-subs', *extra],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1, out.stdout
    return json.loads(lines[0])


def test_reference_arm_prints_the_contract_line():
    sys.path.insert(0, ROOT)
    import bench
    d = _run()
    assert d['impl'] == 'reference'
    assert d['metric'] == bench.metric_name(False) and d['unit'] == 'images/sec' and d['higher_is_better'] is True
    assert d['config']['workload'] == bench.workload_config(False, 1024, 1)['workload']
    assert d['value'] > 0 and d['steps'] == 1 and d['vs_baseline'] is None and d['data'] == 'synthetic'
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == d['value'] and 'batch 32' in cb['sample']
    assert d['e2e'] == {'value': d['value'], 'unit': 'images/sec', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['gpu_launches'] == 0
    # same config object as our arm (the sample description lives outside it), same warm-up rule (W = max(3, --warmup))
    assert d['config'] == bench.workload_config(False, 1024, 1) and d['warmup'] == 3
    # the reference's KAN loop costs per batch: one forward at the labelled batch is reported next to the batch-32 steps
    assert d['same_batch']['batch'] == 1024 and d['same_batch']['value'] > 0


def test_cpu_reference_kan_leg_runs_the_loop_port():
    sys.path.insert(0, ROOT)
    import bench
    r = bench.cpu_reference_kan(batches=(4,), dims=(12, 5, 1))
    assert r['kind'] == 'port' and r['runs'][0]['batch'] == 4 and r['runs'][0]['fwd_bwd_ms'] > r['runs'][0]['fwd_ms'] > 0


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK='1', LOCAL_RANK='1', WORLD_SIZE='2')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1'],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ''


def test_a_failing_sub_benchmark_does_not_take_the_headline_down_on_one_gpu():
    sys.path.insert(0, ROOT)
    import bench
    import pytest

    def boom():
        raise RuntimeError('sub-benchmark failed')
    assert bench.guarded_sub(lambda: {'value': 1.0}, 1) == {'value': 1.0}
    assert 'sub-benchmark failed' in bench.guarded_sub(boom, 1)['error']
    with pytest.raises(RuntimeError):            # several ranks: stay fatal, the peers would hang in the next collective
        bench.guarded_sub(boom, 2)
