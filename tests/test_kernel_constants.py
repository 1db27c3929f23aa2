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


# ------------------------------------------------------------------------------------------ the KAN basis device function on the host
KAN_CU = os.path.join(os.path.dirname(COMMON), 'kan.cu')
HARNESS = r'''
#include <cmath>
#define __device__
#define __forceinline__ inline
namespace {
%s
}
extern "C" void basis_host(const float* t, int n, const float* knots, float* out, float* dout) {
  Knots kn;
  for (int i = 0; i < kKnots; ++i) kn.k[i] = knots[i];
  for (int i = 0; i < n; ++i) {
    float a[kKW], da[kKW];
    const float tc = fminf(fmaxf(t[i], kn.k[0]), kn.k[kKnots - 1]);     // as kan_basis_kernel does (kan.py:16)
    kan_basis_at<true>(tc, kn, a, da);
    for (int k = 0; k < kNB; ++k) { out[i * kNB + k] = a[k]; dout[i * kNB + k] = da[k]; }
  }
}
'''


@pytest.fixture(scope='module')
def basis_host(tmp_path_factory):
    """`kan_basis_at` -- the closed-form truncated cubic B-spline every fp32 KAN kernel evaluates -- cut out of csrc/kan.cu
    VERBATIM (constants + the function) and compiled for the host with g++: the device source itself runs on the CPU."""
    import ctypes
    import subprocess
    src = open(KAN_CU).read()
    cut = src[src.index('constexpr int kNB'):src.index('// the 8 packed activations')]
    assert 'kan_basis_at' in cut and '__global__' not in cut
    d = tmp_path_factory.mktemp('kan_host')
    (d / 'h.cpp').write_text(HARNESS % cut)
    subprocess.run(['g++', '-O1', '-ffp-contract=off', '-shared', '-fPIC', '-o', str(d / 'h.so'), str(d / 'h.cpp')], check=True)
    lib = ctypes.CDLL(str(d / 'h.so'))

    def run(t, knots):
        t = np.ascontiguousarray(t, dtype=F).ravel()
        k = np.ascontiguousarray(knots, dtype=F)
        out, dout = np.zeros((t.size, 7), F), np.zeros((t.size, 7), F)
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        lib.basis_host(p(t), ctypes.c_int(t.size), p(k), p(out), p(dout))
        return out, dout
    return run


def test_device_basis_function_reproduces_the_reference_vectors_on_the_host(basis_host):
    """Against tests/golden/kan_basis.npz = outputs of the reference's own BSplineBasis.compute_basis (models/kan.py:10-44) on
    the knots, their +-1e-6 neighbours, the jump at 0.4 and 4096 random points."""
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'kan_basis.npz'))
    t, want = g['t'].reshape(-1), g['basis'].reshape(-1, 7)
    got, _ = basis_host(t, g['knots'])
    assert np.abs(got - want).max() <= 2e-7
    assert np.array_equal(got == 0, want == 0)                     # same support, incl. identically zero from t >= 0.4
    # SURVEY.md section 4 known answers (RNG-free)
    kat = {-1.0: [0] * 7, -0.9: [.0208333] + [0] * 6, 0.0: [0, 0, .1666667, .6666667, .1666666, 0, 0],
           0.3: [0, 0, 0, .0208333, .4791667, .4791667, .0208333], 0.4: [0] * 7, 1.0: [0] * 7}
    got, _ = basis_host(np.array(list(kat), F), g['knots'])
    assert np.abs(got - np.array(list(kat.values()), F)).max() <= 2e-6


def test_device_basis_derivative_matches_a_finite_difference_on_the_host(basis_host):
    knots = np.linspace(-1, 1, 11).astype(F)
    rng = np.random.default_rng(0)
    t = rng.uniform(-0.99, 0.39, 4000).astype(F)
    t = t[np.abs((t + 1) / 0.2 - np.round((t + 1) / 0.2)) > 0.02]        # stay inside one knot interval for the difference
    h = F(1e-3)
    lo, _ = basis_host(t - h, knots)
    hi, _ = basis_host(t + h, knots)
    _, d = basis_host(t, knots)
    assert np.abs((hi - lo) / (2 * h) - d).max() <= 2e-2 * np.abs(d).max()     # O(h^2) + fp32 cancellation
    _, dead = basis_host(np.array([0.4, 0.7, 1.0], F), knots)
    assert not dead.any()


def test_tiled_residual_stream_layout_is_a_bijection(tmp_path):
    """`xt_offset` / `xt_elem_offset` (common.cuh) place element (row, col) of the fp32 inference residual stream at
    [32-row block][32-col panel][float4 index][lane][4].  Compiled from the header text for the host: every
    (row < R, col < 192) must map to a distinct offset in [0, R * 192) for R a multiple of 32, and a warp (32 consecutive rows)
    reading one float4 index of one panel must touch 512 contiguous bytes."""
    import ctypes
    import subprocess
    text = open(COMMON).read()
    cut = text[text.index('__host__ __device__ __forceinline__ size_t xt_offset'):text.index('#endif  // __CUDACC__')]
    (tmp_path / 'x.cpp').write_text('#include <cstddef>\n#define __host__\n#define __device__\n#define __forceinline__ inline\n' + cut +
                                    '\nextern "C" void offsets(int rows, long long* out) {\n'
                                    '  for (int r = 0; r < rows; ++r) for (int c = 0; c < 192; ++c) out[r * 192 + c] = (long long)xt_elem_offset(r, c);\n}\n'
                                    'extern "C" long long vec_offset(int row, int panel, int j) { return (long long)xt_offset(row, panel, j); }\n')
    subprocess.run(['g++', '-O1', '-shared', '-fPIC', '-o', str(tmp_path / 'x.so'), str(tmp_path / 'x.cpp')], check=True)
    lib = ctypes.CDLL(str(tmp_path / 'x.so'))
    lib.vec_offset.restype = ctypes.c_longlong
    rows = 256
    out = np.zeros(rows * 192, np.int64)
    lib.offsets(rows, out.ctypes.data_as(ctypes.c_void_p))
    assert np.array_equal(np.sort(out), np.arange(rows * 192))
    for panel, j, base_row in ((0, 0, 0), (5, 7, 96), (3, 2, 224)):
        offs = np.array([lib.vec_offset(base_row + lane, panel, j) for lane in range(32)])
        assert np.array_equal(offs - offs[0], 4 * np.arange(32))           # lane l -> floats [4 l, 4 l + 4): 512 contiguous bytes


def test_dropout_rng_is_philox4x32_10(tmp_path):
    """The dropout stream of the heads (`philox4x32_10`, `uniform01` in common.cuh) compiled from the header text for the host
    and checked against the published Random123 known-answer vectors of Philox4x32-10; `uniform01` must be a 24-bit uniform
    in [0, 1) keyed by (seed, offset, element index)."""
    import ctypes
    import subprocess
    text = open(COMMON).read()
    cut = text[text.index('__device__ __forceinline__ uint4 philox4x32_10'):text.index('__device__ __forceinline__ uint32_t smem_u32')]
    shim = ('#include <cstdint>\n#define __device__\n#define __forceinline__ inline\n'
            'struct uint4 { uint32_t x, y, z, w; }; struct uint2 { uint32_t x, y; };\n'
            'static inline uint4 make_uint4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return uint4{a, b, c, d}; }\n'
            'static inline uint2 make_uint2(uint32_t a, uint32_t b) { return uint2{a, b}; }\n'
            'static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }\n')
    (tmp_path / 'p.cpp').write_text(shim + cut.replace('#pragma unroll', '') +
                                    '\nextern "C" void philox(const uint32_t* c, const uint32_t* k, uint32_t* out) {\n'
                                    '  uint4 r = philox4x32_10(make_uint4(c[0], c[1], c[2], c[3]), make_uint2(k[0], k[1]));\n'
                                    '  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;\n}\n'
                                    'extern "C" float uni(unsigned long long s, unsigned long long o, unsigned long long i) { return uniform01(s, o, i); }\n')
    subprocess.run(['g++', '-O1', '-shared', '-fPIC', '-o', str(tmp_path / 'p.so'), str(tmp_path / 'p.cpp')], check=True)
    lib = ctypes.CDLL(str(tmp_path / 'p.so'))
    lib.uni.restype = ctypes.c_float
    lib.uni.argtypes = [ctypes.c_ulonglong] * 3
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        out = (ctypes.c_uint32 * 4)()
        lib.philox((ctypes.c_uint32 * 4)(*ctr), (ctypes.c_uint32 * 2)(*key), out)
        assert tuple(out) == want, [hex(v) for v in out]
    u = np.array([lib.uni(1234, 77, i) for i in range(20000)])
    assert u.min() >= 0.0 and u.max() < 1.0 and np.all(u * 16777216 == np.round(u * 16777216))
    assert abs(u.mean() - 0.5) < 0.01 and abs((u < 0.3).mean() - 0.3) < 0.01           # keep rate of p = 0.3 dropout
    assert lib.uni(1234, 77, 5) != lib.uni(1234, 78, 5) != lib.uni(1235, 77, 5)


# ------------------------------------------------------------------------------------------ the joint-loss device function on the host
HEADS_CU = os.path.join(os.path.dirname(COMMON), 'heads.cu')
LOSS_HARNESS = r'''
#include <cmath>
#include <cstddef>
#include <cstring>
#define __device__
#define __forceinline__ inline
static inline float __int_as_float(int v) { float f; std::memcpy(&f, &v, 4); return f; }
namespace {
%s
}
extern "C" void loss_host(const float* cls, int C, const float* ordl, const float* mu, const float* lv, const float* kan,
                          const long long* yc, const float* ys, const float* alpha, float gamma, int batch, float* sums,
                          float* d_cls, float* d_ord, float* d_mu, float* d_lv, float* d_kan) {
  LossParams p;
  p.cls_logits = cls; p.num_classes = C; p.ord_logits = ordl; p.mu = mu; p.log_var = lv; p.kan = kan;
  p.class_t = yc; p.sev_t = ys; p.alpha = alpha; p.gamma = gamma; p.batch = batch; p.sums = sums;
  p.d_cls = d_cls; p.d_ord = d_ord; p.d_mu = d_mu; p.d_lv = d_lv; p.d_kan = d_kan;
  for (int b = 0; b < batch; ++b) {
    float l[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    joint_loss_sample(p, b, l[0], l[1], l[2], l[3]);
    for (int i = 0; i < 4; ++i) sums[i] += l[i];
  }
}
'''


@pytest.fixture(scope='module')
def loss_host(tmp_path_factory):
    """`LossParams` + `joint_loss_sample` -- the per-sample arithmetic of the fused JointLoss kernel, forward terms and local
    gradients -- cut verbatim out of csrc/heads.cu and compiled for the host."""
    import ctypes
    import subprocess
    src = open(HEADS_CU).read()
    cut = src[src.index('struct LossParams {'):src.index('__global__ void __launch_bounds__(256) joint_loss_kernel')]
    assert 'joint_loss_sample' in cut and '__global__' not in cut
    d = tmp_path_factory.mktemp('loss_host')
    (d / 'l.cpp').write_text(LOSS_HARNESS % cut)
    subprocess.run(['g++', '-O1', '-ffp-contract=off', '-shared', '-fPIC', '-o', str(d / 'l.so'), str(d / 'l.cpp')], check=True)
    lib = ctypes.CDLL(str(d / 'l.so'))

    def run(cls, ordl, mu, lv, kan, yc, ys, alpha, gamma=2.0):
        c = lambda a, dt=F: None if a is None else np.ascontiguousarray(a, dtype=dt)
        cls, ordl, mu, lv, kan, ys, alpha = c(cls), c(ordl), c(mu), c(lv), c(kan), c(ys), c(alpha)
        yc = c(yc, np.int64)
        outs = [np.zeros_like(a) if a is not None else None for a in (cls, ordl, mu, lv, kan)]
        sums = np.zeros(4, F)
        p = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
        lib.loss_host(p(cls), ctypes.c_int(cls.shape[1]), p(ordl), p(mu), p(lv), p(kan), p(yc), p(ys), p(alpha),
                      ctypes.c_float(gamma), ctypes.c_int(cls.shape[0]), p(sums), *[p(o) for o in outs])
        return sums, outs
    return run


def test_joint_loss_device_function_reproduces_the_reference_vectors_on_the_host(loss_host):
    """tests/golden/losses.npz = the reference's JointLoss (training/losses.py:139-181, focal alpha given, gamma 2, weights
    1 / 0.5 / 0.5) and its autograd gradients at stage 4; the stage gating itself is the launcher's (null pointers)."""
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'losses.npz'))
    sums, (d_cls, d_ord, d_mu, d_lv, d_kan) = loss_host(g['in_cls_logits'], g['in_ordinal_logits'], g['in_mu'], g['in_log_var'],
                                                        g['in_kan_severity'], g['yc'], g['ys'], g['alpha'])
    for i, k in enumerate(('cls_loss', 'ord_loss', 'unc_loss', 'kan_loss')):
        assert abs(float(sums[i]) - float(g['s4_' + k])) <= 2e-6 * max(1.0, abs(float(g['s4_' + k]))), k
    total = sums[0] + 1.0 * sums[1] + 0.5 * sums[2] + 0.5 * sums[3]
    assert abs(float(total) - float(g['s4_total_loss'])) <= 5e-6
    for got, w, k in ((d_cls, 1.0, 'cls_logits'), (d_ord, 1.0, 'ordinal_logits'), (d_mu, 0.5, 'mu'), (d_lv, 0.5, 'log_var'),
                      (d_kan, 0.5, 'kan_severity')):
        want = g['s4_d_' + k]
        assert np.abs(w * got - want).max() <= 1e-5 * np.abs(want).max() + 1e-8, k
    # stage 1 = only the focal term reaches the total; its value and gradient do not depend on the other heads
    s1, (d1, *_) = loss_host(g['in_cls_logits'], None, None, None, None, g['yc'], g['ys'], g['alpha'])
    assert abs(float(s1[0]) - float(g['s1_cls_loss'])) <= 2e-6 and not s1[1:].any()
    assert np.abs(d1 - g['s1_d_cls_logits']).max() <= 1e-5 * np.abs(g['s1_d_cls_logits']).max() + 1e-8


def test_joint_loss_device_function_known_answer_and_label_guards(loss_host):
    """the RNG-free known answer of JointLoss(focal_alpha=None), and the label guards: an out-of-range
    class label gives a NaN loss (loud) and a zero gradient instead of an out-of-bounds read; a non-integer focal gamma never
    sees a negative base; fractional severities are compared as floats (`y > k`, losses.py:59)."""
    cls = [[2, .5, -1, 0], [.1, .2, .3, .4]]
    sums, _ = loss_host(cls, [[1, -1, -2], [.5, .5, -.5]], [[.5], [2.5]], [[0], [-1]], [[.3], [2]], [0, 3], [0, 3], None)
    for got, want in zip(sums, (0.328758, 0.612614, -0.017607, 0.545000)):
        assert abs(float(got) - want) <= 2e-6
    assert abs(float(sums[0] + sums[1] + 0.5 * sums[2] + 0.5 * sums[3]) - 1.205068) <= 3e-6
    s, (d, *_) = loss_host(cls, None, None, None, None, [7, 3], [0, 3], None)
    assert np.isnan(s[0]) and not d[0].any() and np.isfinite(d[1]).all() and d[1].any()
    s, (d, *_) = loss_host([[30.0, 0, 0, 0]], None, None, None, None, [0], [0], None, gamma=1.5)       # pt rounds to 1
    assert np.isfinite(s[0]) and np.isfinite(d).all()
    a, _ = loss_host(cls, [[1, -1, -2], [.5, .5, -.5]], None, None, None, [0, 3], [1.5, 0.5], None)
    b, _ = loss_host(cls, [[1, -1, -2], [.5, .5, -.5]], None, None, None, [0, 3], [2.0, 1.0], None)
    assert abs(float(a[1]) - float(b[1])) <= 1e-7             # [1.5 > k] == [2 > k] and [0.5 > k] == [1 > k] for integer k


# ------------------------------------------------------------------------------------------ the tensor-core KAN producer on the host
KAN_TC = os.path.join(os.path.dirname(COMMON), 'kan_tc.cuh')
TC_SHIM = r'''
#include <algorithm>
#include <cfenv>
#include <cmath>
#include <cstdint>
#include <cstring>
#define __device__
#define __forceinline__ inline
#define __global__
#define __restrict__
using std::min;
struct uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return uint4{a, b, c, d}; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline float __fadd_rd(float a, float b) {             // add.rd.f32
  const int old = fegetround();
  fesetround(FE_DOWNWARD);
  volatile float va = a, vb = b;
  volatile float r = va + vb;
  fesetround(old);
  return r;
}
static inline uint16_t bf16_rn(float f) {                       // cvt.rn.bf16.f32 (finite inputs)
  uint32_t u; std::memcpy(&u, &f, 4);
  return static_cast<uint16_t>((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
}
struct __nv_bfloat162 { uint16_t x, y; };                       // .x = low half of the 32-bit word
static inline __nv_bfloat162 __floats2bfloat162_rn(float a, float b) { return __nv_bfloat162{bf16_rn(a), bf16_rn(b)}; }
static inline uint32_t pack_bf16x2(float lo, float hi) { return bf16_rn(lo) | (static_cast<uint32_t>(bf16_rn(hi)) << 16); }
// the two MUFU approximations (ex2.approx, rcp.approx: ~2 ulp) are modelled by the exact operations
static inline float ex2_approx(float x) { return exp2f(x); }
static inline float kan_tc_tanh(float xe, float& rc) {
  const float ex = ex2_approx(xe * 2.8853900817779268f);
  rc = 1.0f / (ex + 1.0f);
  return fmaf(-2.0f, rc, 1.0f);
}
struct Knots { float k[11]; };
struct Dim3 { int x; };
static Dim3 threadIdx;
namespace {
%s
}
extern "C" float host_tanhf(float x) { return tanhf(x); }
extern "C" void thresholds(const float* knots, float* xthr12) {
  Knots kn;
  for (int i = 0; i < 11; ++i) kn.k[i] = knots[i];
  for (int m = 0; m < 9; ++m) { threadIdx.x = m; kan_tc_thresholds_kernel(kn, xthr12); }
  xthr12[9] = xthr12[10] = xthr12[11] = INFINITY;
}
extern "C" void expand(const float* x, int n, const float* xthr12, uint8_t* hi, uint8_t* lo) {
  for (int i = 0; i < n; ++i) kan_tc_expand_store(x[i], 16u * i, hi, lo, xthr12);
}
'''


@pytest.fixture(scope='module')
def tc_producer(tmp_path_factory):
    """`kan_tc_thresholds_kernel`, `kan_tc_interval` and `kan_tc_expand_store` -- the operand producer of the three tcgen05 KAN
    kernels: interval by a round-down add, calibrated x-space thresholds next to a knot, Horner cubics, bf16 hi / lo split, one-hot
    placement as a 128-bit shift -- cut verbatim out of csrc/kan_tc.cuh and compiled for the host (CUDA intrinsics shimmed)."""
    import ctypes
    import subprocess
    src = open(KAN_TC).read()
    a = src.index('__global__ void kan_tc_thresholds_kernel')
    cut = src[a:src.index('// WpT (fp32', a)]
    b = src.index('__device__ __forceinline__ void kan_tc_interval')
    cut += src[b:src.index('// tanh x = 1 - 2 /', b)]
    c = src.index('__device__ __forceinline__ void kan_tc_expand_store')
    cut += src[c:src.index('__global__ void __launch_bounds__(kTcThreads, 1)', c)]
    d = tmp_path_factory.mktemp('tc_host')
    (d / 't.cpp').write_text(TC_SHIM % cut)
    subprocess.run(['g++', '-O1', '-frounding-math', '-fno-strict-aliasing', '-ffp-contract=off', '-shared', '-fPIC', '-o',
                    str(d / 't.so'), str(d / 't.cpp')], check=True)
    lib = ctypes.CDLL(str(d / 't.so'))
    lib.host_tanhf.restype = ctypes.c_float
    lib.host_tanhf.argtypes = [ctypes.c_float]
    # torch.linspace(-1, 1, 11) in fp32 (knots[5] = -1.49e-8, knots[7] = 0.39999998): numpy's linspace rounds differently
    knots = np.load(os.path.join(ROOT, 'tests', 'golden', 'kan_basis.npz'))['knots'].astype(F)
    xthr = np.zeros(12, F)
    p = lambda arr: arr.ctypes.data_as(ctypes.c_void_p)
    lib.thresholds(p(knots), p(xthr))

    def expand(x):
        x = np.ascontiguousarray(x, dtype=F)
        hi, lo = np.zeros((x.size, 8), np.uint16), np.zeros((x.size, 8), np.uint16)
        lib.expand(p(x), ctypes.c_int(x.size), p(xthr), p(hi), p(lo))
        f = lambda h: (h.astype(np.uint32) << 16).view(F)
        return f(hi), f(lo)
    return lib, knots, xthr, expand


def test_tensor_core_kan_producer_on_the_host(tc_producer, basis_host):
    lib, knots, xthr, expand = tc_producer
    assert knots.shape == (11,) and abs(float(knots[7]) - 0.4) < 1e-7
    # thresholds: xthr[m] is the smallest float whose tanhf reaches knot m
    for m in range(1, 8):
        assert lib.host_tanhf(float(xthr[m])) >= knots[m] > lib.host_tanhf(float(np.nextafter(xthr[m], F(-np.inf))))
    assert xthr[0] == -np.inf and np.all(np.isinf(xthr[8:]))
    rng = np.random.default_rng(0)
    near = np.concatenate([[np.nextafter(F(xthr[m]), F(s * np.inf)) if k else F(xthr[m]) for k in (0, 1, 2) for s in (-1, 1)]
                           for m in range(1, 8)]).astype(F)
    for m in range(1, 8):                                   # +-2 ulps as well
        near = np.concatenate([near, [np.nextafter(np.nextafter(F(xthr[m]), F(np.inf)), F(np.inf)),
                                      np.nextafter(np.nextafter(F(xthr[m]), F(-np.inf)), F(-np.inf))]]).astype(F)
    x = np.concatenate([rng.normal(0, 1.5, 20000), rng.uniform(-0.01, 0.01, 500), near, [0.0, -0.0, 12.0, -12.0, 40.0, -40.0]]).astype(F)
    hi, lo = expand(x)
    got = hi.astype(np.float64) + lo.astype(np.float64)
    # expected: the reference basis at t = tanhf(x) through the verbatim fp32 device function (itself pinned to the reference's
    # vectors above), and the raw input in slot 7
    t = np.array([lib.host_tanhf(float(v)) for v in x], F)
    want, _ = basis_host(t, knots)
    assert np.abs(got[:, :7] - want).max() <= 1e-5              # measured 4.1e-6: bf16 hi + lo = 16 mantissa bits, fast tanh 3e-7 in t
    assert np.abs(got[:, 7] - x).max() <= 2.0 ** -16 * np.abs(x).max()
    # identical interval decisions, also ON a threshold and one / two ulps either side: nothing outside the live window j-3 .. j,
    # nothing at all in the dead zone tanh x >= 0.4
    j = (t[:, None] >= knots[None, 1:]).sum(1)
    slots = np.arange(7)[None, :]
    outside = (slots > j[:, None]) | (slots < j[:, None] - 3) | (j[:, None] >= 7)
    assert not hi[:, :7][outside].any() and not lo[:, :7][outside].any()
    inside = ~outside & (want > 1e-6)
    assert np.all(hi[:, :7][inside] > 0)
    dead = t >= knots[7]
    assert dead.sum() > 3000 and not got[dead, :7].any()
    assert np.abs(got[~dead, :7].sum(1) - want[~dead].sum(1)).max() <= 2e-5


def test_small_kan_segment_function_and_activations_on_the_host(tmp_path, basis_host):
    """`kan_segment` (csrc/kan_small.cuh: the four live cubic segments without the one-hot placement -- what the small-layer
    kernels and the fused heads + KAN tail evaluate) against `kan_basis_at`, itself pinned to the reference's vectors above; and
    `act_grad` (the activation derivative recovered from the layer OUTPUT, kan.cu) against a finite difference of `kan_act`
    (ReLU between the KAN layers, 3 * sigmoid on the last one: models/kan.py:138-149)."""
    import ctypes
    import subprocess
    kan = open(KAN_CU).read()
    small = open(os.path.join(os.path.dirname(COMMON), 'kan_small.cuh')).read()
    consts = kan[kan.index('constexpr int kNB'):kan.index('// basis values (and optionally d/dt)')]
    seg = small[small.index('template <bool DERIV>\n__device__ __forceinline__ bool kan_segment'):small.index('// sW[(i*8 + k) * NOUT + o]')]
    act = small[small.index('__device__ __forceinline__ float kan_act'):small.index('template <int NOUT>\n__global__')]
    agrad = kan[kan.index('__device__ __forceinline__ float act_grad'):kan.index('// ------------------------------------------------------------------ weight packing')]
    (tmp_path / 's.cpp').write_text('#include <cmath>\n#define __device__\n#define __forceinline__ inline\nnamespace {\n' + consts + seg + act + agrad + '}\n'
                                    'extern "C" int segment(float t, const float* knots, float* v, float* d) {\n'
                                    '  Knots kn; for (int i = 0; i < kKnots; ++i) kn.k[i] = knots[i];\n'
                                    '  int j = -1; float vv[4] = {0, 0, 0, 0}, dd[4] = {0, 0, 0, 0};\n'
                                    '  const bool live = kan_segment<true>(t, kn, j, vv, dd);\n'
                                    '  for (int m = 0; m < 4; ++m) { v[m] = vv[m]; d[m] = dd[m]; }\n'
                                    '  return live ? j : -1;\n}\n'
                                    'extern "C" float act_fwd(int a, float v) { return kan_act(a, v); }\n'
                                    'extern "C" float act_bwd(int a, float y) { return act_grad(a, y); }\n')
    subprocess.run(['g++', '-O1', '-ffp-contract=off', '-shared', '-fPIC', '-o', str(tmp_path / 's.so'), str(tmp_path / 's.cpp')], check=True)
    lib = ctypes.CDLL(str(tmp_path / 's.so'))
    lib.act_fwd.restype = lib.act_bwd.restype = ctypes.c_float
    lib.act_fwd.argtypes = lib.act_bwd.argtypes = [ctypes.c_int, ctypes.c_float]
    lib.segment.argtypes = [ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'kan_basis.npz'))
    knots = g['knots'].astype(F)
    t = np.clip(g['t'].reshape(-1), knots[0], knots[-1]).astype(F)
    want, dwant = basis_host(t, knots)
    got, dgot = np.zeros_like(want), np.zeros_like(dwant)
    v, d = np.zeros(4, F), np.zeros(4, F)
    for n, tv in enumerate(t):
        j = lib.segment(float(tv), knots.ctypes.data_as(ctypes.c_void_p), v.ctypes.data_as(ctypes.c_void_p), d.ctypes.data_as(ctypes.c_void_p))
        for m in range(4):
            if j >= 0 and 0 <= j - m < 7:
                got[n, j - m], dgot[n, j - m] = v[m], d[m]
    assert np.abs(got - want).max() <= 2e-7 and np.abs(dgot - dwant).max() <= 2e-5 * np.abs(dwant).max()
    assert np.array_equal(got == 0, want == 0)
    for a in (0, 1, 2):
        for z in (-3.0, -0.7, 0.3, 1.9, 4.0):
            y = lib.act_fwd(a, z)
            fd = (lib.act_fwd(a, z + 1e-2) - lib.act_fwd(a, z - 1e-2)) / 2e-2
            assert abs(lib.act_bwd(a, y) - fd) <= 2e-3, (a, z)
    assert lib.act_fwd(2, 0.0) == 1.5 and lib.act_fwd(1, -1.0) == 0.0 and lib.act_bwd(1, 0.0) == 0.0


def test_predict_decode_kernel_on_the_host(tmp_path):
    """`predict_decode_kernel` (csrc/heads.cu: the epilogue of RoViTKAN.predict, rovit_kan.py:126-161 with heads.py:45-77) run
    thread by thread on the host against the oracle's restatement: softmax, first-maximum argmax, the ordinal decode
    p0 = c0, pk = ck - ck-1, p3 = 1 - c2 (negative entries included), its expected value, std = exp(log_var / 2)."""
    import ctypes
    import subprocess
    import sys
    import torch
    sys.path.insert(0, ROOT)
    from oracle import heads as oheads
    src = open(HEADS_CU).read()
    a = src.index('__global__ void predict_decode_kernel')
    cut = src[a:src.index('// out = {cls, ord, unc, kan', a)]
    (tmp_path / 'd.cpp').write_text('#include <cmath>\n#include <cstddef>\n#define __global__\n#define __restrict__\n'
                                    'struct D3 { int x; }; static D3 blockIdx, blockDim, threadIdx;\n' + cut +
                                    'extern "C" void decode(const float* cls, int C, const float* ordl, const float* lv, int batch, long long* idx,\n'
                                    '                       float* probs, float* oprobs, float* osev, float* std_) {\n'
                                    '  blockIdx.x = 0; blockDim.x = batch + 3;\n'
                                    '  for (int b = 0; b < batch + 3; ++b) { threadIdx.x = b; predict_decode_kernel(cls, C, ordl, lv, batch, idx, probs, oprobs, osev, std_); }\n}\n')
    subprocess.run(['g++', '-O1', '-ffp-contract=off', '-shared', '-fPIC', '-o', str(tmp_path / 'd.so'), str(tmp_path / 'd.cpp')], check=True)
    lib = ctypes.CDLL(str(tmp_path / 'd.so'))
    rng = np.random.default_rng(3)
    B, C = 257, 4
    cls = rng.normal(0, 2, (B, C)).astype(F)
    cls[5] = [1.0, 3.0, 3.0, -1.0]                                # a tie: torch.argmax returns the first maximum
    ordl = rng.normal(0, 2, (B, C - 1)).astype(F)                 # unordered cumulative logits -> some negative "probabilities"
    lv = rng.uniform(-10, 10, (B, 1)).astype(F)
    idx, probs, oprobs = np.full(B, -1, np.int64), np.zeros((B, C), F), np.zeros((B, C), F)
    osev, std = np.zeros((B, 1), F), np.zeros((B, 1), F)
    p = lambda arr: arr.ctypes.data_as(ctypes.c_void_p)
    lib.decode(p(cls), ctypes.c_int(C), p(ordl), p(lv), ctypes.c_int(B), p(idx), p(probs), p(oprobs), p(osev), p(std))
    tc, to, tl = torch.from_numpy(cls), torch.from_numpy(ordl), torch.from_numpy(lv)
    want_probs = torch.softmax(tc, dim=1)
    assert np.abs(probs - want_probs.numpy()).max() <= 2e-7
    assert np.array_equal(idx, torch.argmax(want_probs, dim=1).numpy()) and idx[5] == 1
    want_op = oheads.ordinal_probabilities(to).numpy()
    assert np.abs(oprobs - want_op).max() <= 2e-7 and (want_op < 0).any()
    assert np.abs(osev - oheads.ordinal_severity(to).numpy()).max() <= 1e-6
    assert np.abs(std - torch.exp(0.5 * tl).numpy()).max() <= 1e-6 * float(torch.exp(0.5 * tl).max())


# ------------------------------------------------------------------------------------------ the fused optimizer tail on the host
OPT_SHIM = r'''
#include <cmath>
#include <cstddef>
#include <cstdint>
#define __global__
#define __launch_bounds__(n)
#define __grid_constant__
#define __shared__ static
#define __restrict__
#define __device__
#define __forceinline__ inline
using std::isfinite;
struct float4 { float x, y, z, w; };
struct D3 { unsigned x; };
static D3 threadIdx, blockIdx;
static inline void __syncthreads() {}
static inline float __shfl_xor_sync(unsigned, float v, int) { return v; }      // (the norm kernel is compiled, never run here)
static inline void atomicAdd(float* p, float v) { *p += v; }
namespace {
%s
}
// table + hyper-parameters exactly as rvk_optimizer_step_impl fills them (optimizer.cu), then the update kernel block by block,
// thread by thread (thread 0 first: it publishes the block's tensor index), then the finish kernel
extern "C" void opt_step(int n, float** params, const float** grads, const long long* numel, const int* group, float* exp_avg,
                         float* exp_avg_sq, float* state4, const double* lr, int n_groups, double beta1, double beta2, double eps,
                         double weight_decay, float max_grad_norm, float grad_mult, const float* grad_scale, const float* found_inf) {
  OptHyper H{};
  for (int i = 0; i < n_groups; ++i) {
    H.lr[i] = static_cast<float>(lr[i]); H.lr_d[i] = lr[i];
    H.decay[i] = static_cast<float>(1.0 - lr[i] * weight_decay);
  }
  H.om_beta1 = static_cast<float>(1.0 - beta1); H.om_beta2 = static_cast<float>(1.0 - beta2);
  H.beta1_d = beta1; H.beta2_d = beta2;
  H.beta1 = static_cast<float>(beta1); H.beta2 = static_cast<float>(beta2); H.eps = static_cast<float>(eps);
  H.weight_decay = static_cast<float>(weight_decay); H.max_norm = max_grad_norm;
  H.grad_mult = grad_mult; H.grad_scale = grad_scale; H.found_inf = found_inf;
  OptTable T{};
  T.n = n;
  int chunks = 0;
  for (int i = 0; i < n; ++i) {
    T.p[i] = params[i]; T.g[i] = grads[i]; T.chunk_start[i] = chunks; T.numel[i] = static_cast<int>(numel[i]);
    T.group[i] = static_cast<unsigned char>(group[i]);
    chunks += chunks_of(numel[i]);
  }
  T.chunk_start[n] = chunks;
  for (int b = 0; b < chunks; ++b)
    for (int t = 0; t < kOptThreads; ++t) { blockIdx.x = b; threadIdx.x = t; optim_update_kernel(T, H, exp_avg, exp_avg_sq, state4, 0); }
  blockIdx.x = 0; threadIdx.x = 0;
  optim_finish_kernel(H, state4);
}
extern "C" long long padded(long long numel) { return static_cast<long long>(chunks_of(numel)) * kChunk; }
'''


def test_fused_adamw_update_kernel_matches_torch_on_the_host(tmp_path):
    """`optim_update_kernel` + `optim_finish_kernel` (csrc/optimizer.cu: GradScaler unscale + inf check + global-norm clip + AdamW
    with the reference's two LR groups, training/trainer.py:118-129, training/optimizer.py:18-25) run on the host against
    `clip_grad_norm_` + `torch.optim.AdamW` for three steps; the sum of squares the norm kernel would deliver is supplied."""
    import ctypes
    import subprocess
    import torch
    src = open(os.path.join(os.path.dirname(COMMON), 'optimizer.cu')).read()
    cut = src[src.index('constexpr int kChunk'):src.index('}  // namespace')]
    (tmp_path / 'o.cpp').write_text(OPT_SHIM % cut)
    subprocess.run(['g++', '-O1', '-ffp-contract=off', '-shared', '-fPIC', '-o', str(tmp_path / 'o.so'), str(tmp_path / 'o.cpp')], check=True)
    lib = ctypes.CDLL(str(tmp_path / 'o.so'))
    lib.padded.restype = ctypes.c_longlong
    lib.padded.argtypes = [ctypes.c_longlong]
    torch.manual_seed(0)
    sizes, groups, lrs = [5000, 17, 4096, 300, 1], [0, 0, 1, 1, 1], [1e-5, 1e-4]
    ref = [torch.nn.Parameter(torch.randn(s)) for s in sizes]
    opt = torch.optim.AdamW([{'params': ref[:2], 'lr': lrs[0]}, {'params': ref[2:], 'lr': lrs[1]}], weight_decay=1e-4, foreach=False)
    ours = [p.detach().clone().numpy() for p in ref]
    offs = np.cumsum([0] + [lib.padded(s) for s in sizes])
    m, v, state = np.zeros(offs[-1], F), np.zeros(offs[-1], F), np.zeros(4, F)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)

    def step(grads, scale=None, found_inf=None, max_norm=1.0):
        total = np.float32(sum(float((g.astype(np.float64) ** 2).sum()) for g in grads))
        state[0] = total
        P = (ctypes.c_void_p * 5)(*[a.ctypes.data for a in ours])
        G = (ctypes.c_void_p * 5)(*[g.ctypes.data for g in grads])
        lib.opt_step(5, P, G, vp(np.array(sizes, np.int64)), vp(np.array(groups, np.int32)), vp(m), vp(v), vp(state),
                     vp(np.array(lrs, np.float64)), 2, ctypes.c_double(0.9), ctypes.c_double(0.999), ctypes.c_double(1e-8),
                     ctypes.c_double(1e-4), ctypes.c_float(max_norm), ctypes.c_float(1.0),
                     None if scale is None else vp(scale), None if found_inf is None else vp(found_inf))

    for it, gscale in enumerate((0.002, 0.05, 0.0005)):               # clipping active in the second step only
        grads = [(torch.randn(s) * gscale) for s in sizes]
        for p, g in zip(ref, grads):
            p.grad = g.clone()
        norm = torch.nn.utils.clip_grad_norm_(ref, 1.0)
        opt.step()
        step([g.numpy().copy() for g in grads])
        assert state[1] == it + 1 and state[3] == 1.0 and state[0] == 0.0
        assert abs(float(state[2]) - float(norm)) <= 2e-6 * float(norm)
        assert (float(norm) > 1.0) == (it == 1)
        for a, p, o in zip(ours, ref, offs):
            assert np.abs(a - p.detach().numpy()).max() <= 2e-6 * float(p.detach().abs().max()), it
            st = opt.state[p]
            assert np.abs(m[o:o + a.size] - st['exp_avg'].numpy()).max() <= 2e-6 * float(st['exp_avg'].abs().max()) + 1e-12
            assert np.abs(v[o:o + a.size] - st['exp_avg_sq'].numpy()).max() <= 2e-6 * float(st['exp_avg_sq'].abs().max()) + 1e-20
    # GradScaler path: gradients arrive multiplied by the loss scale; a non-finite gradient or found_inf skips the step entirely
    before = [a.copy() for a in ours]
    grads = [np.random.default_rng(1).normal(0, 0.01, s).astype(F) for s in sizes]
    bad = [g.copy() for g in grads]
    bad[2][7] = np.inf
    step(bad)
    assert state[3] == 0.0 and state[1] == 3 and all(np.array_equal(a, b) for a, b in zip(ours, before))
    step(grads, found_inf=np.ones(1, F))
    assert state[3] == 0.0 and state[1] == 3 and all(np.array_equal(a, b) for a, b in zip(ours, before))
    saved = ([a.copy() for a in ours], m.copy(), v.copy())
    step([g * np.float32(1024.0) for g in grads], scale=np.full(1, 1024.0, F), found_inf=np.zeros(1, F))
    assert state[3] == 1.0 and state[1] == 4
    from_scaled = ([a.copy() for a in ours], m.copy(), v.copy())
    for a, b in zip(ours, saved[0]):
        a[...] = b
    m[...], v[...], state[1] = saved[1], saved[2], 3
    step(grads)                               # the same step from unscaled gradients: 1024 is a power of two, the unscale is exact
    assert all(np.array_equal(a, b) for a, b in zip(ours, from_scaled[0]))
    assert np.array_equal(m, from_scaled[1]) and np.array_equal(v, from_scaled[2]) and not all(np.array_equal(a, b) for a, b in zip(ours, before))


def test_patch_gather_kernel_on_the_host(tmp_path):
    """`im2col_kernel<FMT>` (csrc/encoder_kernels.cu: timm's PatchEmbed Conv2d(3,192,16,16) as a gather, fp32 / bf16 / uint8 pixels)
    run thread by thread on the host: column = c*256 + ky*16 + kx (the flattened Conv2d weight), token 0 an all-zero row, fp32
    pixels rounded to nearest-even bf16 (bit-exact against torch), uint8 pixels normalised as pixel * 1/(255 std) - mean/std
    (the reference's ToTensor + Normalize folded in: within one bf16 ulp of the transform, exact against the folded form)."""
    import ctypes
    import subprocess
    import torch
    src = open(os.path.join(os.path.dirname(COMMON), 'encoder_kernels.cu')).read()
    consts = src[src.index('constexpr int kD = 192;'):src.index('// ------------------------------------------------------------------ patch extraction')]
    a = src.index('struct PixelNorm')
    cut = consts + src[a:src.index('// table[0] = cls_token', a)]
    shim = ('#include <cmath>\n#include <cstddef>\n#include <cstdint>\n#include <cstring>\n#define __global__\n#define __restrict__\n'
            'struct uint4 { uint32_t x, y, z, w; }; struct uint2 { uint32_t x, y; }; struct float4 { float x, y, z, w; };\n'
            'struct __nv_bfloat16 { uint16_t v; };\n'
            'static inline uint4 make_uint4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return uint4{a, b, c, d}; }\n'
            'static inline uint16_t bf16_rn(float f) { uint32_t u; std::memcpy(&u, &f, 4); return (uint16_t)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16); }\n'
            'static inline uint32_t pack_bf16x2(float lo, float hi) { return bf16_rn(lo) | ((uint32_t)bf16_rn(hi) << 16); }\n'
            'struct D3 { unsigned x; }; static D3 threadIdx, blockIdx, blockDim;\n')
    run = ('\ntemplate <int FMT> static void run(const void* img, uint16_t* out, int batch, const float* n6) {\n'
           '  PixelNorm nrm{}; for (int i = 0; i < 3; ++i) { nrm.scale[i] = n6[i]; nrm.shift[i] = n6[3 + i]; }\n'
           '  const long long threads = (long long)batch * kTok * 3 * 32;\n'
           '  blockDim.x = 256;\n'
           '  for (long long t = 0; t < threads + 64; ++t) { blockIdx.x = (unsigned)(t / 256); threadIdx.x = (unsigned)(t % 256);\n'
           '    im2col_kernel<FMT>(img, reinterpret_cast<__nv_bfloat16*>(out), batch, nrm); }\n}\n'
           'extern "C" void gather(int fmt, const void* img, uint16_t* out, int batch, const float* n6) {\n'
           '  if (fmt == 0) run<0>(img, out, batch, n6); else if (fmt == 1) run<1>(img, out, batch, n6); else run<2>(img, out, batch, n6);\n}\n')
    (tmp_path / 'g.cpp').write_text(shim + cut + run)
    subprocess.run(['g++', '-O1', '-ffp-contract=off', '-fno-strict-aliasing', '-shared', '-fPIC', '-o', str(tmp_path / 'g.so'), str(tmp_path / 'g.cpp')], check=True)
    lib = ctypes.CDLL(str(tmp_path / 'g.so'))
    vp = lambda arr: arr.ctypes.data_as(ctypes.c_void_p)
    B = 2
    g = torch.Generator().manual_seed(0)
    img = torch.randn(B, 3, 224, 224, generator=g)

    def unfold(x):          # (B,3,224,224) -> (B,197,768), column c*256 + ky*16 + kx, row 0 zero
        p = x.reshape(B, 3, 14, 16, 14, 16).permute(0, 2, 4, 1, 3, 5).reshape(B, 196, 768)
        return torch.cat([torch.zeros(B, 1, 768, dtype=x.dtype), p], dim=1)
    bits = lambda t: t.to(torch.bfloat16).view(torch.int16).numpy().astype(np.uint16)
    zero6 = np.zeros(6, F)
    out = np.full((B, 197, 768), 0xffff, np.uint16)
    lib.gather(0, vp(img.numpy()), vp(out), B, vp(zero6))
    assert np.array_equal(out, bits(unfold(img)))                                          # fp32 pixels: bit-exact
    out1 = np.full_like(out, 0xffff)
    img16 = np.ascontiguousarray(bits(img))
    lib.gather(1, vp(img16), vp(out1), B, vp(zero6))
    assert np.array_equal(out1, out)                                                       # bf16 pixels: identical result
    u8 = torch.randint(0, 256, (B, 3, 224, 224), generator=g, dtype=torch.uint8)
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    n6 = np.array([1.0 / (255.0 * s) for s in std] + [-m / s for m, s in zip(mean, std)], F)
    out2 = np.full_like(out, 0xffff)
    lib.gather(2, vp(np.ascontiguousarray(u8.numpy())), vp(out2), B, vp(n6))
    folded = torch.from_numpy(np.float32(u8.numpy()) * n6[:3].reshape(1, 3, 1, 1) + n6[3:].reshape(1, 3, 1, 1))   # one rounding less than fmaf
    normal = (u8.float() / 255.0 - torch.tensor(mean).view(1, 3, 1, 1)) / torch.tensor(std).view(1, 3, 1, 1)
    got = torch.from_numpy(out2.astype(np.int16)).view(torch.bfloat16).float()
    assert float((got - unfold(normal)).abs().max()) <= 2.0 ** -7 * 2.7                   # one bf16 ulp at |x| <= 2.7
    assert float((got - unfold(folded)).abs().max()) <= 2.0 ** -7 * 2.7 and float((got - unfold(folded)).abs().mean()) < 2e-3
    assert not out2[:, 0].any() and not out[:, 0].any()
