"""Weighted minimax fit behind common.cuh::gelu_erf_grad:  gelu'(x) = x >= 0 ? 1 - D(|x|) : D(|x|),
D(a) = Phi(-a) - a phi(a) = exp(-a^2 / 2) * R(a);  R as a degree-6 polynomial on [0, 5.5], error weighted by exp(-a^2 / 2)
(Lawson iterations), then checked in emulated fp32 Horner arithmetic against the closed form."""
import numpy as np
from numpy.polynomial import chebyshev as C, polynomial as Pl
from scipy.special import ndtr

t = np.linspace(0, 5.5, 40001)
w = np.exp(-t * t / 2)
R = (ndtr(-t) - t * w / np.sqrt(2 * np.pi)) / w
V = C.chebvander(t / 5.5 * 2 - 1, 6)
lw = np.ones_like(t)
for _ in range(200):
    coef = np.linalg.lstsq(V * (w * lw)[:, None], R * w * lw, rcond=None)[0]
    err = np.abs((V @ coef - R) * w)
    lw = lw * (err / err.max() + 1e-3) ** 0.5
    lw /= lw.mean()
c = Pl.Polynomial(C.cheb2poly(coef))(Pl.Polynomial([-1.0, 2 / 5.5])).coef
print('fit error', err.max())
print('coefficients, low -> high:', ['%.9e' % v for v in c])
c32 = c.astype(np.float32)
z = np.linspace(-8, 8, 200001).astype(np.float32)
a = np.minimum(np.abs(z), np.float32(5.5))
p = np.full_like(a, c32[6])
for k in range(5, -1, -1):
    p = (p * a + c32[k]).astype(np.float32)
d = (np.exp2((a * np.float32(-0.72134752044448170368) * a).astype(np.float32)).astype(np.float32) * p).astype(np.float32)
g = np.where(z >= 0, np.float32(1) - d, d)
exact = ndtr(z.astype(np.float64)) + z * np.exp(-z.astype(np.float64) ** 2 / 2) / np.sqrt(2 * np.pi)
print('max |gelu_grad - exact| in fp32 Horner form:', np.abs(g - exact).max())
