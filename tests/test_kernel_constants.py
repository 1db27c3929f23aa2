"""CPU checks of the numerical constants baked into the kernels (no GPU, no CUDA call): the polynomial coefficients are parsed
out of csrc/common.cuh and the device functions are re-evaluated in emulated fp32 Horner arithmetic against closed forms.

  gelu_erf       = relu(x) - |x| * 2^P(|x|),  P = degree-6 fit of log2(Phi(-a))       (timm nn.GELU, approximate='none')
  gelu_erf_grad  = x >= 0 ? 1 - D : D,  D = 2^(-a^2 log2(e)/2) * R(a), R degree 6       (backward of the same)

`ex2.approx.ftz.f32` (2 ulp) is modelled by an exact exp2 rounded to fp32, so the bounds asserted here are those the header
comments state plus a small allowance; an edited or mistyped coefficient moves the error by orders of magnitude."""

import os
import re

import numpy as np
import pytest
from scipy.special import ndtr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMMON = os.path.join(ROOT, 'rovit-kan-interpretable-vision-transformer-for-rose-disease-severity-estimation_b200', 'csrc', 'common.cuh')
F = np.float32


def horner_coefficients(func_name: str, var: str):
    """Coefficients high -> low of the `var = fmaf(c6, t, c5); var = fmaf(var, t, c4); ...` chain inside `func_name`."""
    text = open(COMMON).read()
    body = text[text.index(f'float {func_name}(float'):]
    body = body[:body.index('\n}\n')]
    first = re.search(rf'float {var} = fmaf\(([-+0-9.e]+)f, t, ([-+0-9.e]+)f\);', body)
    rest = re.findall(rf'{var} = fmaf\({var}, t, ([-+0-9.e]+)f\);', body)
    assert first and len(rest) == 5, (func_name, first, rest)
    return [F(first.group(1)), F(first.group(2))] + [F(c) for c in rest]


def horner(coefs, t):
    p = np.full_like(t, coefs[0])
    for c in coefs[1:]:
        p = (p * t + c).astype(F)          # fmaf: one rounding, which float64-product-then-round reproduces for fp32 inputs
    return p


@pytest.fixture(scope='module')
def grid():
    return np.concatenate([np.linspace(-9, 9, 400001), [0.0, -0.0, 5.5, -5.5, 1e-6, -1e-6]]).astype(F)


def test_gelu_forward_polynomial(grid):
    c = horner_coefficients('norm_cdf_neg', 'p')
    a = np.minimum(np.abs(grid), F(5.5))
    phi_neg = np.exp2(horner(c, a).astype(np.float64)).astype(F)
    got = (np.maximum(grid, F(0)) - np.abs(grid) * phi_neg).astype(F)
    x = grid.astype(np.float64)
    exact = x * ndtr(x)
    err = np.abs(got - exact)
    assert err.max() <= 6e-6, err.max()                                  # header: <= 4e-6 absolute
    inside = np.abs(x) < 5.5
    rel = err[inside] / np.maximum(np.abs(exact[inside]), 1e-30)
    assert rel[np.abs(exact[inside]) > 1e-6].max() <= 5e-5               # header: <= 2.7e-5 relative for |x| < 5.5
    assert got[np.where(grid == 0)[0]].max() == 0.0                      # gelu(+-0) = 0 exactly (ReLU gates downstream compare with 0)


def test_gelu_backward_polynomial(grid):
    c = horner_coefficients('gelu_erf_grad', 'r')
    a = np.minimum(np.abs(grid), F(5.5))
    d = (np.exp2((a * F(-0.72134752044448170368) * a).astype(F).astype(np.float64)).astype(F) * horner(c, a)).astype(F)
    got = np.where(grid >= 0, F(1) - d, d)
    x = grid.astype(np.float64)
    exact = ndtr(x) + x * np.exp(-x * x / 2) / np.sqrt(2 * np.pi)
    assert np.abs(got - exact).max() <= 2.5e-5                           # header / tools/fit_gelu_grad.py: 1.6e-5
    assert abs(float(got[grid == 0][0]) - 0.5) <= 2e-5 and got.min() >= -0.13 and got.max() <= 1.13   # range of gelu'


def test_gelu_backward_is_the_derivative_of_the_forward_polynomial(grid):
    """Consistency of the two fits with each other: a central difference of the forward form reproduces the backward form
    (float64 evaluation of the same coefficients, so only the fits' own errors remain)."""
    cf = [float(v) for v in horner_coefficients('norm_cdf_neg', 'p')]
    cb = [float(v) for v in horner_coefficients('gelu_erf_grad', 'r')]
    x = np.linspace(-5, 5, 20001)

    def fwd(v):
        a = np.minimum(np.abs(v), 5.5)
        return np.maximum(v, 0) - np.abs(v) * np.exp2(np.polyval(cf, a))
    a = np.abs(x)
    d = np.exp2(-0.72134752044448170368 * a * a) * np.polyval(cb, a)
    bwd = np.where(x >= 0, 1 - d, d)
    h = 1e-4
    assert np.abs((fwd(x + h) - fwd(x - h)) / (2 * h) - bwd).max() <= 2e-4
