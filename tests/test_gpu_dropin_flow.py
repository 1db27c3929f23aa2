"""GPU side of the drop-in boundary: a train + checkpoint + evaluate run written against the REFERENCE'S module names
(tests/dropin_flow_script.py), started through `python -m rovitkan_b200.launch` like a reference script would be."""

import json
import math
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_train_checkpoint_evaluate_through_the_hook(tmp_path):
    env = dict(os.environ, PYTHONPATH=ROOT, ROVITKAN_SYNTH_PER_CLASS='10')
    r = subprocess.run([sys.executable, '-m', 'rovitkan_b200.launch', os.path.join(ROOT, 'tests', 'dropin_flow_script.py'),
                        '--output_dir', str(tmp_path), '--epochs', '2', '--batch_size', '8'], cwd=str(tmp_path), env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith('FLOW_RESULT ')][-1]
    res = json.loads(line[len('FLOW_RESULT '):])
    assert len(res['history']) == 2 and res['n_test'] == 40 and res['reload_identical']
    for h in res['history']:
        assert math.isfinite(h['train_loss']) and math.isfinite(h['val_loss'])
    assert res['predict_keys'] == sorted(['class', 'class_probs', 'features', 'ordinal_probs', 'ordinal_severity',
                                          'uncertainty_mu', 'uncertainty_std', 'kan_severity'])
    assert 0.0 <= res['sev_range'][0] <= res['sev_range'][1] <= 3.0
    assert 'Backbone frozen' in r.stdout and 'Backbone unfrozen' in r.stdout
