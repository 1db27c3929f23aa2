import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with -m gpu')


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def T(a):
    return torch.from_numpy(np.asarray(a))


@pytest.fixture(scope='session')
def golden():
    return load_golden


def assert_close(a, b, rtol=1e-3, atol=1e-5, what='', scale_tol=0.0):
    """|a-b| <= rtol*|b| + atol + scale_tol*max|b| elementwise, with a readable failure.
    `scale_tol` is for tensors with large dynamic range (gradients), where fp32
    re-association noise is proportional to the largest entries."""
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    assert a.shape == b.shape, f'{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}'
    err = (a - b).abs()
    tol = rtol * b.abs() + atol + scale_tol * (float(b.abs().max()) if b.numel() else 0.0)
    bad = err > tol
    if bad.any():
        i = int((err - tol).argmax())
        raise AssertionError(
            f'{what}: {int(bad.sum())}/{bad.numel()} out of tolerance (rtol={rtol}, atol={atol}); '
            f'worst at flat {i}: got {a.flatten()[i].item():.8g} want {b.flatten()[i].item():.8g} '
            f'(max abs err {err.max().item():.3g})')
