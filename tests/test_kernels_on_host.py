"""Whole CUDA kernels on the CPU (no GPU): the kernel text is cut VERBATIM out of the .cu sources and compiled by g++ against
tests/host_emu/cuda_host_emu.h, a thread-level emulation (every CUDA thread of a block a fiber scheduled round-robin between
barriers: real __syncthreads / warp-shuffle / shared-memory / atomicAdd semantics, deterministic), then run with the launch
geometry the library uses and compared with the reference-generated vectors, the oracle or PyTorch on the same inputs.

Covered: the CUDA-core kernels -- the fp32 KAN layer kernels and the small-output KAN kernels (forward and every gradient), the
fused heads + KAN tail in inference and in training (forward, backward, all 23 parameter gradients, dropout), the per-layer linear
path, the joint-loss kernel with its block reduction, LayerNorm forward / backward, the optimizer's gradient-norm kernel, attention
probabilities, weight shadows, token table, column sums.  The single-CTA tcgen05 GEMMs (gemm_nt in every epilogue mode, gemm_tn) run under a
FUNCTIONAL emulation of mbarriers / TMA / tensor memory / tcgen05.mma (tests/host_emu/tcgen05_host_emu.h).  Not emulated: the
kernels with TMEM-resident operands and CTA pairs (attention, the fused MLP block) and the tensor-core KAN kernels (their
CUDA-core operand producer is covered in tests/test_kernel_constants.py) -- those are tested on the GPU (`-m gpu`)."""

import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, 'rovit-kan-interpretable-vision-transformer-for-rose-disease-severity-estimation_b200', 'csrc')
EMU = os.path.join(ROOT, 'tests', 'host_emu')
F = np.float32


def read(name):
    return open(os.path.join(CSRC, name)).read()


def between(text, start, end):
    a = text.index(start)
    return text[a:text.index(end, a)]


def compile_host(tmp, name, body):
    src = tmp / f'{name}.cpp'
    src.write_text('#include "cuda_host_emu.h"\n' + body)
    so = tmp / f'{name}.so'
    # ROVITKAN_EMU_SANITIZE=1 (with libasan preloaded into the interpreter, see test_kernels_under_address_sanitizer): every global /
    # shared / workspace access of the emulated kernels is bounds-checked
    san = ['-fsanitize=address,undefined', '-fno-sanitize-recover=undefined', '-fno-omit-frame-pointer'] if os.environ.get('ROVITKAN_EMU_SANITIZE') == '1' else []
    r = subprocess.run(['g++', '-O1', '-std=c++17', '-ffp-contract=off', '-fno-strict-aliasing', '-I', EMU, '-shared', '-fPIC', *san,
                        '-o', str(so), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    return ctypes.CDLL(str(so))


def vp(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


COMMON_BITS = None


def common_bits():
    global COMMON_BITS
    if COMMON_BITS is None:
        c = read('common.cuh')
        COMMON_BITS = (between(c, '__device__ __forceinline__ float warp_sum', '__device__ __forceinline__ uint32_t pack_bf16x2')
                       + between(c, '__host__ __device__ __forceinline__ size_t xt_offset', '#endif  // __CUDACC__'))
    return COMMON_BITS


# ------------------------------------------------------------------------------------------ LayerNorm
@pytest.fixture(scope='module')
def layernorm_lib(tmp_path_factory):
    k = read('encoder_kernels.cu')
    body = ('namespace {\n' + common_bits() + between(k, 'constexpr int kD = 192;', '// ------------------------------------------------------------------ patch extraction')
            + between(k, '// ------------------------------------------------------------------ LayerNorm forward',
                      '// ------------------------------------------------------------------ column sums') + '}\n' + r'''
extern "C" void ln_fwd(int out_bf16, const float* x, long long xs, const float* gamma, const float* beta, float eps, void* y,
                       long long ys, float* mean, float* rstd, int rows, int grid, int block) {
  EmuDim g; g.x = grid; EmuDim b; b.x = block;
  if (out_bf16) emu_launch(g, b, 0, [=] { layernorm_fwd_kernel<true, false>(x, xs, gamma, beta, eps, y, ys, mean, rstd, rows); });
  else emu_launch(g, b, 0, [=] { layernorm_fwd_kernel<false, false>(x, xs, gamma, beta, eps, y, ys, mean, rstd, rows); });
}
extern "C" void ln_bwd(const float* g_, long long gs, const float* x, long long xs, const float* mean, const float* rstd,
                       const float* gamma, const float* dx_in, float* dx_out, long long dxs, uint16_t* dx_bf16, float* dgamma,
                       float* dbeta, float* dcolsum, int rows, int grid) {
  EmuDim g; g.x = grid; EmuDim b; b.x = 256;
  emu_launch(g, b, 0, [=] { layernorm_bwd_kernel<false>(g_, gs, x, xs, mean, rstd, gamma, dx_in, dx_out, dxs,
                                                        reinterpret_cast<__nv_bfloat16*>(dx_bf16), dgamma, dbeta, dcolsum, rows); });
}
''')
    lib = compile_host(tmp_path_factory.mktemp('ln'), 'ln', body)
    P, I, L, Fl = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float
    lib.ln_fwd.argtypes = [I, P, L, P, P, Fl, P, L, P, P, I, I, I]
    lib.ln_bwd.argtypes = [P, L, P, L, P, P, P, P, P, L, P, P, P, P, I, I]
    return lib


def test_layernorm_forward_kernel_on_the_host(layernorm_lib):
    """`layernorm_fwd_kernel` (warp per 192-wide row, two shuffle reductions; timm's LayerNorm, eps 1e-6) incl. the strided-row
    form that pools the CLS rows (row stride 197 * 192) and the bf16 output that feeds the qkv GEMM."""
    torch.manual_seed(0)
    rows = 37
    x = torch.randn(rows, 192) * 2 + 0.5
    gamma, beta = torch.randn(192), torch.randn(192)
    want = torch.nn.functional.layer_norm(x, (192,), gamma, beta, 1e-6)
    y, mean, rstd = np.zeros((rows, 192), F), np.zeros(rows, F), np.zeros(rows, F)
    layernorm_lib.ln_fwd(0, vp(x.numpy()), 192, vp(gamma.numpy()), vp(beta.numpy()), 1e-6, vp(y), 192, vp(mean), vp(rstd),
                         rows, 3, 128)
    assert np.abs(y - want.numpy()).max() <= 2e-6 * float(want.abs().max())
    assert np.abs(mean - x.mean(1).numpy()).max() <= 1e-6 and np.abs(rstd - (x.var(1, unbiased=False) + 1e-6).rsqrt().numpy()).max() <= 1e-5
    yb = np.zeros((rows, 192), np.uint16)
    layernorm_lib.ln_fwd(1, vp(x.numpy()), 192, vp(gamma.numpy()), vp(beta.numpy()), 1e-6, vp(yb), 192, None, None, rows, 2, 256)
    got = torch.from_numpy(yb.astype(np.int16)).view(torch.bfloat16).float()
    assert float((got - want).abs().max()) <= 2.0 ** -8 * float(want.abs().max())
    # strided rows: every 5th row of a taller matrix (the CLS pooling reads row b * 197 of the token stream)
    tall = torch.randn(5 * 9, 192)
    y2 = np.zeros((9, 192), F)
    layernorm_lib.ln_fwd(0, vp(tall.numpy()), 5 * 192, vp(gamma.numpy()), vp(beta.numpy()), 1e-6, vp(y2), 192, None, None, 9, 1, 64)
    assert np.abs(y2 - torch.nn.functional.layer_norm(tall[::5], (192,), gamma, beta, 1e-6).numpy()).max() <= 1e-5


@pytest.mark.parametrize('rows,grid,with_dx_in', [(50, 2, True), (16, 1, False), (131, 3, True)])
def test_layernorm_backward_kernel_on_the_host(layernorm_lib, rows, grid, with_dx_in):
    """`layernorm_bwd_kernel` (16 lanes per row, half-warp shuffles, per-warp shared partials, one atomicAdd per column and
    block): dx (+ the residual gradient dx_in), its bf16 copy, d gamma, d beta and the fused column sums of dx (= the bias
    gradient of the Linear layer in front) against autograd."""
    torch.manual_seed(rows)
    x = (torch.randn(rows, 192) * 1.5).requires_grad_(True)
    gamma = torch.randn(192, requires_grad=True)
    beta = torch.zeros(192, requires_grad=True)
    g = torch.randn(rows, 192)
    dx_in = torch.randn(rows, 192) if with_dx_in else None
    torch.nn.functional.layer_norm(x, (192,), gamma, beta, 1e-6).backward(g)
    want_dx = x.grad + (dx_in if with_dx_in else 0)
    xd = x.detach()
    mean = xd.mean(1).numpy().copy()
    rstd = (xd.var(1, unbiased=False) + 1e-6).rsqrt().numpy().copy()
    dx, dxb = np.zeros((rows, 192), F), np.zeros((rows, 192), np.uint16)
    dgamma, dbeta, dcol = np.zeros(192, F), np.zeros(192, F), np.zeros(192, F)
    layernorm_lib.ln_bwd(vp(g.numpy()), 192, vp(xd.numpy()), 192, vp(mean), vp(rstd), vp(gamma.detach().numpy()),
                         vp(dx_in.numpy()) if with_dx_in else None, vp(dx), 192, vp(dxb), vp(dgamma), vp(dbeta), vp(dcol), rows, grid)
    tol = lambda t: 3e-6 * float(t.abs().max())
    assert np.abs(dx - want_dx.numpy()).max() <= tol(want_dx)
    assert np.abs(dgamma - gamma.grad.numpy()).max() <= 2e-5 * float(gamma.grad.abs().max())
    assert np.abs(dbeta - beta.grad.numpy()).max() <= 2e-5 * float(beta.grad.abs().max())
    assert np.abs(dcol - want_dx.sum(0).numpy()).max() <= 2e-5 * float(want_dx.sum(0).abs().max())
    got_b = torch.from_numpy(dxb.astype(np.int16)).view(torch.bfloat16).float()
    assert float((got_b - want_dx).abs().max()) <= 2.0 ** -8 * float(want_dx.abs().max())


# ------------------------------------------------------------------------------------------ the fp32 KAN layer kernels
@pytest.fixture(scope='module')
def kan_lib(tmp_path_factory):
    k = read('kan.cu')
    body = ('namespace {\n' + between(k, 'constexpr int kNB = 7;', 'int pad_to(int v, int m)') + '}\n' + r'''
static int pad_to_(int v, int m) { return (v + m - 1) / m * m; }
// the launch sequence of rvk_kan_layer_fwd_launch / rvk_kan_layer_bwd_launch (kan.cu) for the fp32 CUDA-core path
extern "C" void kan_layer(const float* x, const float* spline, const float* lin_w, const float* lin_b, const float* knots, int batch,
                          int n_in, int n_out, int act, int spt, float* y, const float* gy, float* dx, float* dspline, float* dlin_w,
                          float* dlin_b) {
  Knots kn;
  for (int i = 0; i < kKnots; ++i) kn.k[i] = knots[i];
  const int in_pad = pad_to_(n_in, 16), out_pad = pad_to_(n_out, 64);
  const long long wp = static_cast<long long>(in_pad) * 8 * out_pad;
  std::vector<float> ws(3 * wp, -7.0f);                 // workspaces come uninitialised
  float *Wp = ws.data(), *WpT = ws.data() + wp, *dWp = ws.data() + 2 * wp;
  EmuDim b256; b256.x = 256;
  EmuDim g; g.x = static_cast<unsigned>(std::min<long long>((wp + 255) / 256, 1184));
  emu_launch(g, b256, 0, [=] { kan_pack_kernel(spline, lin_w, n_in, n_out, in_pad, out_pad, Wp, WpT); });
  EmuDim gf; gf.y = out_pad / kTO;
  if (spt == 1) { gf.x = (batch + 15) / 16; emu_launch(gf, b256, 0, [=] { kan_fwd_kernel<1>(x, Wp, lin_b, kn, y, act, batch, n_in, n_out, in_pad, out_pad); }); }
  else { gf.x = (batch + kTS - 1) / kTS; emu_launch(gf, b256, 0, [=] { kan_fwd_kernel<4>(x, Wp, lin_b, kn, y, act, batch, n_in, n_out, in_pad, out_pad); }); }
  if (gy == nullptr) return;
  std::fill(dWp, dWp + wp, 0.0f);                        // cudaMemsetAsync
  const int tiles = (in_pad / kIC) * (out_pad / kTO);
  int splits = (2 * 148 + tiles - 1) / tiles;
  const int sub_tiles = (batch + kTS - 1) / kTS;
  if (splits > sub_tiles) splits = sub_tiles;
  const int sps = ((sub_tiles + splits - 1) / splits) * kTS;
  splits = (batch + sps - 1) / sps;
  EmuDim gw; gw.x = in_pad / kIC; gw.y = out_pad / kTO; gw.z = splits;
  emu_launch(gw, b256, 0, [=] { kan_bwd_w_kernel(x, y, gy, act, kn, dWp, dlin_b, batch, n_in, n_out, out_pad, sps); });
  const long long tot = static_cast<long long>(n_in) * n_out * 8;
  EmuDim gu; gu.x = static_cast<unsigned>(std::min<long long>((tot + 255) / 256, 1184));
  emu_launch(gu, b256, 0, [=] { kan_unpack_grad_kernel(dWp, n_in, n_out, out_pad, dspline, dlin_w); });
  EmuDim gx; gx.x = (batch + kTS - 1) / kTS; gx.y = (n_in + kDxIC - 1) / kDxIC;
  emu_launch(gx, b256, 0, [=] { kan_bwd_x_kernel(x, y, gy, act, kn, WpT, dx, batch, n_in, n_out, in_pad, out_pad); });
}
''')
    lib = compile_host(tmp_path_factory.mktemp('kan'), 'kan', body)
    P, I = ctypes.c_void_p, ctypes.c_int
    lib.kan_layer.argtypes = [P, P, P, P, P, I, I, I, I, I, P, P, P, P, P, P]
    return lib


@pytest.mark.parametrize('tag,spt', [('l0', 1), ('l1', 1), ('l2', 4), ('odd', 1), ('odd', 4)])
def test_kan_layer_kernels_reproduce_the_reference_on_the_host(kan_lib, tag, spt):
    """The fused KAN layer as the library launches it below batch 8192 -- `kan_pack_kernel` -> `kan_fwd_kernel<SPT>`; backward
    `kan_bwd_w_kernel` -> `kan_unpack_grad_kernel`, `kan_bwd_x_kernel` -- against tests/golden/kan_layers.npz: outputs and autograd
    gradients of the reference's own KANLayer (models/kan.py:70-95: tanh, truncated basis, per-pair loop, linear branch on the raw
    input) for 192->64, 64->16, 16->1 and a ragged 10->3 layer at batch 7."""
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'kan_layers.npz'))
    c = lambda k: np.ascontiguousarray(g[f'{tag}_{k}'], dtype=F)
    x, sw, lw, lb, gy, knots = c('x'), c('sw'), c('lw'), c('lb'), c('gy'), c('knots')
    batch, n_in = x.shape
    n_out = lw.shape[0]
    y, dx = np.full((batch, n_out), np.nan, F), np.full((batch, n_in), np.nan, F)
    dsw, dlw, dlb = np.zeros_like(sw), np.zeros_like(lw), np.zeros_like(lb)          # gradients accumulate (+=)
    kan_lib.kan_layer(vp(x), vp(sw), vp(lw), vp(lb), vp(knots), batch, n_in, n_out, 0, spt, vp(y), vp(gy), vp(dx), vp(dsw), vp(dlw), vp(dlb))
    for got, key in ((y, 'y'), (dx, 'dx'), (dsw, 'dsw'), (dlw, 'dlw'), (dlb, 'dlb')):
        want = c(key)
        assert np.isfinite(got).all(), key
        assert np.abs(got - want).max() <= 2e-5 * np.abs(want).max() + 1e-6, (key, float(np.abs(got - want).max()), float(np.abs(want).max()))


# ------------------------------------------------------------------------------------------ the small-output KAN kernels
@pytest.fixture(scope='module')
def kan_small_lib(tmp_path_factory):
    k, s = read('kan.cu'), read('kan_small.cuh')
    cut = between(s, 'constexpr int kSmThreads', '// few outputs: always')
    dyn = 'extern __shared__ __align__(16) float sm_small[];'
    assert cut.count(dyn) == 2
    cut = cut.replace(dyn, 'float* sm_small = static_cast<float*>(emu_dynamic_smem());')
    body = ('namespace {\n' + between(k, 'constexpr int kNB = 7;', '// basis values (and optionally d/dt)')
            + between(k, '__device__ __forceinline__ float act_grad', '// ------------------------------------------------------------------ weight packing')
            + cut + '}\n' + r'''
// launch geometry of kan_small_fwd_launch_t / kan_small_bwd_launch_t (kan_small.cuh), 148 SMs
template <int NOUT>
static void run_small(const float* x, const float* spline, const float* lin_w, const float* lin_b, const Knots& kn, int batch, int n_in,
                      int n_out, int act, float* y, const float* gy, float* dx, float* dspline, float* dlin_w, float* dlin_b) {
  EmuDim blk; blk.x = kSmThreads;
  {
    const int warps = kSmThreads / 32;
    int grid = (batch + warps - 1) / warps;
    if (grid > 148 * 4) grid = 148 * 4;
    EmuDim g; g.x = grid;
    emu_launch(g, blk, static_cast<size_t>(n_in) * kKW * NOUT * 4,
               [=] { kan_small_fwd_kernel<NOUT>(x, spline, lin_w, lin_b, kn, y, act, batch, n_in, n_out); });
  }
  if (gy == nullptr) return;
  constexpr int OQ = NOUT <= 4 ? 1 : NOUT / 4;
  constexpr int OPT = NOUT < 4 ? NOUT : 4;
  const int G = kSmThreads / (n_in * OQ);
  int ctas = 148 * 4;
  const int min_per_cta = G * 8;
  if (static_cast<long long>(ctas) * min_per_cta > batch) ctas = (batch + min_per_cta - 1) / min_per_cta;
  if (ctas < 1) ctas = 1;
  const int spc = (batch + ctas - 1) / ctas;
  ctas = (batch + spc - 1) / spc;
  EmuDim g; g.x = ctas;
  emu_launch(g, blk, static_cast<size_t>(n_in * kKW * NOUT + (kKW * OPT + OPT) * kSmThreads) * 4,
             [=] { kan_small_bwd_kernel<NOUT>(x, y, gy, spline, lin_w, kn, act, dx, dspline, dlin_w, dlin_b, batch, n_in, n_out, spc); });
}
extern "C" void kan_small(const float* x, const float* spline, const float* lin_w, const float* lin_b, const float* knots, int batch,
                          int n_in, int n_out, int act, float* y, const float* gy, float* dx, float* dspline, float* dlin_w,
                          float* dlin_b) {
  Knots kn;
  for (int i = 0; i < kKnots; ++i) kn.k[i] = knots[i];
#define RUN(N) run_small<N>(x, spline, lin_w, lin_b, kn, batch, n_in, n_out, act, y, gy, dx, dspline, dlin_w, dlin_b)
  if (n_out <= 1) RUN(1); else if (n_out <= 2) RUN(2); else if (n_out <= 4) RUN(4); else if (n_out <= 8) RUN(8); else RUN(16);
}
''')
    lib = compile_host(tmp_path_factory.mktemp('kan_small'), 'kan_small', body)
    P, I = ctypes.c_void_p, ctypes.c_int
    lib.kan_small.argtypes = [P, P, P, P, P, I, I, I, I, P, P, P, P, P, P]
    return lib


@pytest.mark.parametrize('tag', ['l1', 'l2', 'odd'])
def test_small_kan_kernels_reproduce_the_reference_on_the_host(kan_small_lib, tag):
    """`kan_small_fwd_kernel` / `kan_small_bwd_kernel` (the kernels the library actually launches for the 64->16 and 16->1 layers
    of the production stack and the 64->1 layer of the microbenchmark: warp per sample forward; ONE backward kernel for dx, dW,
    dWl, db with shared-memory accumulators, shuffles between output quads and a per-CTA flush) against the reference's KANLayer."""
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'kan_layers.npz'))
    c = lambda k: np.ascontiguousarray(g[f'{tag}_{k}'], dtype=F)
    x, sw, lw, lb, gy, knots = c('x'), c('sw'), c('lw'), c('lb'), c('gy'), c('knots')
    batch, n_in = x.shape
    n_out = lw.shape[0]
    y, dx = np.full((batch, n_out), np.nan, F), np.full((batch, n_in), np.nan, F)
    dsw, dlw, dlb = np.zeros_like(sw), np.zeros_like(lw), np.zeros_like(lb)
    kan_small_lib.kan_small(vp(x), vp(sw), vp(lw), vp(lb), vp(knots), batch, n_in, n_out, 0, vp(y), vp(gy), vp(dx), vp(dsw), vp(dlw), vp(dlb))
    for got, key in ((y, 'y'), (dx, 'dx'), (dsw, 'dsw'), (dlw, 'dlw'), (dlb, 'dlb')):
        want = c(key)
        assert np.isfinite(got).all(), key
        assert np.abs(got - want).max() <= 2e-5 * np.abs(want).max() + 1e-6, (key, float(np.abs(got - want).max()), float(np.abs(want).max()))


def test_kan_stack_on_the_host_matches_the_reference_module(kan_lib, kan_small_lib):
    """KANSeverityModule (models/kan.py:138-149) as the library chains it below batch 8192: 192->64 through the general kernels
    with the ReLU fused (act 1), 64->16 (ReLU) and 16->1 (3 * sigmoid, act 2) through the small-output kernels; forward and the
    input gradient through all three layers against the reference module's output / autograd (golden `mod_*`, seed 11)."""
    import sys
    sys.path.insert(0, ROOT)
    from oracle import kan as okan
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'kan_layers.npz'))
    x, gy, want_y, want_dx = (np.ascontiguousarray(g[k], dtype=F) for k in ('mod_x', 'mod_gy', 'mod_y', 'mod_dx'))
    torch.manual_seed(11)                                   # the module's parameters are re-derived from the seed of make_golden.py
    layers = []
    for a, b in ((192, 64), (64, 16), (16, 1)):
        sw = torch.randn(a, b, 7) * 0.1
        lin = torch.nn.Linear(a, b)
        layers.append((sw.numpy().copy(), lin.weight.detach().numpy().copy(), lin.bias.detach().numpy().copy()))
    knots = np.ascontiguousarray(g['l0_knots'], dtype=F)
    check = okan.severity_forward(torch.from_numpy(x), [tuple(torch.from_numpy(t) for t in l) for l in layers], torch.from_numpy(knots))
    if float((check - torch.from_numpy(want_y)).abs().max()) > 1e-5:
        pytest.skip('torch RNG stream differs from the one the golden file was made with')
    acts, ys, cur = (1, 1, 2), [], x
    for li, (sw, lw, lb) in enumerate(layers):
        y = np.full((x.shape[0], lw.shape[0]), np.nan, F)
        if li == 0:
            kan_lib.kan_layer(vp(cur), vp(sw), vp(lw), vp(lb), vp(knots), x.shape[0], lw.shape[1], lw.shape[0], acts[li], 1, vp(y), None, None, None, None, None)
        else:
            kan_small_lib.kan_small(vp(cur), vp(sw), vp(lw), vp(lb), vp(knots), x.shape[0], lw.shape[1], lw.shape[0], acts[li], vp(y), None, None, None, None, None)
        ys.append(y)
        cur = y
    assert np.abs(ys[-1] - want_y).max() <= 1e-5 and ys[-1].min() >= 0.0 and ys[-1].max() <= 3.0
    grad, ins = gy, [x] + ys[:-1]
    for li in (2, 1, 0):
        sw, lw, lb = layers[li]
        dx = np.full_like(ins[li], np.nan)
        dsw, dlw, dlb, yy = np.zeros_like(sw), np.zeros_like(lw), np.zeros_like(lb), ys[li].copy()
        fn = kan_lib.kan_layer if li == 0 else kan_small_lib.kan_small
        args = [vp(ins[li]), vp(sw), vp(lw), vp(lb), vp(knots), x.shape[0], lw.shape[1], lw.shape[0], acts[li]] + ([1] if li == 0 else [])
        fn(*args, vp(yy), vp(np.ascontiguousarray(grad)), vp(dx), vp(dsw), vp(dlw), vp(dlb))
        grad = dx
    assert np.abs(grad - want_dx).max() <= 2e-5 * np.abs(want_dx).max() + 1e-7


# ------------------------------------------------------------------------------------------ the fused multi-task tail (north_star (c))
HEAD_KEYS = ['classification_head.fc1.weight', 'classification_head.fc1.bias', 'classification_head.fc2.weight', 'classification_head.fc2.bias',
             'ordinal_head.fc1.weight', 'ordinal_head.fc1.bias', 'ordinal_head.fc2.weight', 'ordinal_head.fc2.bias',
             'uncertainty_head.fc1.weight', 'uncertainty_head.fc1.bias', 'uncertainty_head.fc_mu.weight', 'uncertainty_head.fc_mu.bias',
             'uncertainty_head.fc_logvar.weight', 'uncertainty_head.fc_logvar.bias'] + [
    f'kan_module.kan_layers.{l}.{n}' for l in range(3) for n in ('spline_weights', 'linear.weight', 'linear.bias')]


@pytest.fixture(scope='module')
def tail_lib(tmp_path_factory):
    import re
    k, c = read('kan.cu'), read('common.cuh')
    hf, ht = read('heads_fused.cuh'), read('heads_train.cuh')
    cp = between(hf, '__device__ __forceinline__ void hf_cp_async16', '// TRAIN: the same kernel')
    hf = hf.replace(cp, 'static inline void hf_cp_async16(float* smem_dst, const float* gsrc) { std::memcpy(smem_dst, gsrc, 16); }   // cp.async 16 B\n\n')
    dyn = 'extern __shared__ __align__(16) float hsm[];'
    both = (hf + ht).replace('#pragma once', '')
    assert both.count(dyn) == 2 and both.count('asm volatile("cp.async') == 4
    both = both.replace(dyn, 'float* hsm = static_cast<float*>(emu_dynamic_smem());')
    both = re.sub(r'asm volatile\("cp\.async\.(commit_group|wait_group 0);" ::: "memory"\);', ';', both)
    assert 'asm' not in both
    body = ('namespace {\n' + between(c, '__device__ __forceinline__ uint4 philox4x32_10', '__device__ __forceinline__ uint32_t smem_u32')
            + between(k, 'constexpr int kNB = 7;', '// ------------------------------------------------------------------ weight packing')
            + both + '}\n' + r'''
// the launch sequences of rvk_heads_fused_prepare / rvk_heads_fused / rvk_heads_train_forward / rvk_heads_train_backward (kan.cu)
static HeadsFusedParams params_of(const float* const* p23) {
  HeadsFusedParams p;
  auto Fp = [&](int i) { return p23[i]; };
  p.fc1_w[0] = Fp(0); p.fc1_b[0] = Fp(1); p.fc2_w[0] = Fp(2); p.fc2_b[0] = Fp(3);
  p.fc1_w[1] = Fp(4); p.fc1_b[1] = Fp(5); p.fc2_w[1] = Fp(6); p.fc2_b[1] = Fp(7);
  p.fc1_w[2] = Fp(8); p.fc1_b[2] = Fp(9); p.fc2_w[2] = Fp(10); p.fc2_b[2] = Fp(11); p.fc2_w[3] = Fp(12); p.fc2_b[3] = Fp(13);
  for (int l = 0; l < 3; ++l) { p.spline[l] = Fp(14 + 3 * l); p.lin_w[l] = Fp(15 + 3 * l); p.lin_b[l] = Fp(16 + 3 * l); }
  return p;
}
extern "C" int tail_ws_floats() { return kHfWsFloats; }
extern "C" void tail_forward(const float* const* p23, const float* knots, const float* feat, int batch, int train, float drop_p,
                             unsigned long long seed, unsigned long long offset, float* ws, float* cls, float* ord, float* mu,
                             float* lv, float* kan, float* h, float* a1, float* a2) {
  Knots kn;
  for (int i = 0; i < kKnots; ++i) kn.k[i] = knots[i];
  const HeadsFusedParams p = params_of(p23);
  EmuDim g; g.x = (kHfWsFloats + 255) / 256; EmuDim b; b.x = 256;
  emu_launch(g, b, 0, [=] { heads_fused_pack_kernel(p, ws); });
  EmuDim gt; gt.x = (batch + kHfS - 1) / kHfS; EmuDim bt; bt.x = kHfThreads;
  if (train) {
    HeadsTrainSave sv{h, a1, a2, drop_p, seed, offset};
    emu_launch(gt, bt, kHfSmemBytes, [=] { heads_fused_kernel<true>(feat, ws, kn, batch, cls, ord, mu, lv, kan, sv); });
  } else {
    emu_launch(gt, bt, kHfSmemBytes, [=] { heads_fused_kernel<false>(feat, ws, kn, batch, cls, ord, mu, lv, kan, HeadsTrainSave{}); });
  }
}
extern "C" void tail_backward(const float* knots, const float* feat, const float* ws, int batch, float drop_p, const float* h,
                              const float* a1, const float* a2, const float* lv, const float* kan, const float* d_cls, const float* d_ord,
                              const float* d_mu, const float* d_lv, const float* d_kan, float* dfeat, float* dws, float* const* g23) {
  Knots kn;
  for (int i = 0; i < kKnots; ++i) kn.k[i] = knots[i];
  std::fill(dws, dws + kHfWsFloats, 0.0f);               // cudaMemsetAsync
  HeadsTrainBwdArgs a{feat, ws, h, a1, a2, lv, kan, d_cls, d_ord, d_mu, d_lv, d_kan, dfeat, dws, drop_p, batch};
  EmuDim gt; gt.x = (batch + kHfS - 1) / kHfS; EmuDim bt; bt.x = kHtThreads;
  emu_launch(gt, bt, kHtSmemBytes, [=] { heads_train_bwd_kernel(a, kn); });
  HeadsGradPtrs gp;
  auto G = [&](int i) { return g23[i]; };
  gp.fc1_w[0] = G(0); gp.fc1_b[0] = G(1); gp.fc2_w[0] = G(2); gp.fc2_b[0] = G(3);
  gp.fc1_w[1] = G(4); gp.fc1_b[1] = G(5); gp.fc2_w[1] = G(6); gp.fc2_b[1] = G(7);
  gp.fc1_w[2] = G(8); gp.fc1_b[2] = G(9); gp.fc2_w[2] = G(10); gp.fc2_b[2] = G(11); gp.fc2_w[3] = G(12); gp.fc2_b[3] = G(13);
  for (int l = 0; l < 3; ++l) { gp.spline[l] = G(14 + 3 * l); gp.lin_w[l] = G(15 + 3 * l); gp.lin_b[l] = G(16 + 3 * l); }
  EmuDim g; g.x = (kHfWsFloats + 255) / 256; EmuDim b; b.x = 256;
  emu_launch(g, b, 0, [=] { heads_fused_unpack_grad_kernel(dws, gp); });
}
''')
    lib = compile_host(tmp_path_factory.mktemp('tail'), 'tail', body)
    P, I, Fl, U = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_ulonglong
    lib.tail_forward.argtypes = [P, P, P, I, I, Fl, U, U, P, P, P, P, P, P, P, P, P]
    lib.tail_backward.argtypes = [P, P, P, I, Fl, P, P, P, P, P, P, P, P, P, P, P, P, P]
    return lib


def _tail_setup(batch, seed=0):
    import sys
    sys.path.insert(0, ROOT)
    from oracle import model as omodel
    sd = {k: v.clone() for k, v in omodel.random_state_dict(0).items() if not k.startswith('backbone')}
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():          # timm-style init leaves the head biases at their defaults; move everything off zero, and drive
        for key in HEAD_KEYS:      # some log-variances into the clamp at +-10 (heads.py:99)
            if key.endswith('bias'):
                sd[key] += torch.randn(sd[key].shape, generator=g) * 0.05
        sd['uncertainty_head.fc_logvar.weight'] *= 60.0
    feat = torch.randn(batch, 192, generator=g) * 0.8
    return omodel, sd, feat


def _run_tail_forward(lib, sd, feat, train, drop_p=0.0, seed=0, offset=0):
    batch = feat.shape[0]
    params = [np.ascontiguousarray(sd[k].numpy(), dtype=F) for k in HEAD_KEYS]
    table = (ctypes.c_void_p * 23)(*[p.ctypes.data for p in params])
    knots = np.ascontiguousarray(sd['kan_module.kan_layers.0.knots'].numpy(), dtype=F)
    ws = np.full(lib.tail_ws_floats(), np.nan, F)
    f = np.ascontiguousarray(feat.numpy(), dtype=F)
    mk = lambda n: np.full((batch, n), np.nan, F)
    out = {'cls': mk(4), 'ord': mk(3), 'mu': mk(1), 'lv': mk(1), 'kan': mk(1), 'h': mk(384), 'a1': mk(64), 'a2': mk(16)}
    lib.tail_forward(table, vp(knots), vp(f), batch, int(train), drop_p, seed, offset, vp(ws), *[vp(out[k]) for k in ('cls', 'ord', 'mu', 'lv', 'kan', 'h', 'a1', 'a2')])
    return params, knots, ws, f, out


@pytest.mark.parametrize('batch', [13, 8])
def test_fused_inference_tail_kernel_on_the_host(tail_lib, batch):
    """`heads_fused_pack_kernel` + `heads_fused_kernel<false>`: the three MLP heads and the KAN stack 192->64->16->1 of
    RoViTKAN.forward (rovit_kan.py:96-124) in ONE kernel -- 512 threads x 8 samples per CTA, split-K partial sums through shared
    memory, cp.async-prefetched late weights, 16- and 32-lane shuffle reductions -- against the oracle (itself pinned to the
    reference's heads / KAN modules); batch 13 leaves a ragged last CTA."""
    omodel, sd, feat = _tail_setup(batch)
    _, _, _, _, out = _run_tail_forward(tail_lib, sd, feat, train=False)
    with torch.no_grad():
        ref = omodel.heads_forward(sd, feat, 4)
    for k, rk in (('cls', 'cls_logits'), ('ord', 'ordinal_logits'), ('mu', 'mu'), ('lv', 'log_var'), ('kan', 'kan_severity')):
        want = ref[rk].numpy()
        assert np.abs(out[k] - want).max() <= 2e-5 * max(1.0, float(np.abs(want).max())), k
    assert (np.abs(out['lv']) == 10.0).any() and (np.abs(out['lv']) < 10.0).any()          # both sides of the clamp are exercised
    assert np.array_equal(out['cls'].argmax(1), ref['cls_logits'].numpy().argmax(1))
    assert np.array_equal((out['ord'] > 0).sum(1), (ref['ordinal_logits'].numpy() > 0).sum(1))


def test_fused_training_tail_kernels_on_the_host(tail_lib):
    """north_star (c) in training: `heads_fused_kernel<true>` (forward, saves the hidden activations) and `heads_train_bwd_kernel`
    + `heads_fused_unpack_grad_kernel` (ONE backward kernel for d features and all 23 parameter gradients, atomics into a packed
    buffer) against torch autograd through the oracle: outputs, d features and every parameter gradient, clamp zero-gradient
    included; then with Dropout p = 0.3: the keep mask recovered from the saved activations is a Philox mask of the right rate,
    and outputs and gradients equal the oracle's under that very mask (heads.py:20,41,94)."""
    batch = 11
    omodel, sd, feat = _tail_setup(batch, seed=5)
    from oracle import heads as oheads
    from oracle import kan as okan
    gen = torch.Generator().manual_seed(9)
    ups = {k: torch.randn(batch, n, generator=gen) for k, n in (('cls', 4), ('ord', 3), ('mu', 1), ('lv', 1), ('kan', 1))}

    def oracle(keep3=None):
        sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k in HEAD_KEYS}
        sdg['kan_module.kan_layers.0.knots'] = sd['kan_module.kan_layers.0.knots']
        f = feat.clone().requires_grad_(True)
        w = lambda k: sdg[k]
        kc, ko, ku = keep3 if keep3 is not None else (None, None, None)
        cls = oheads.classification_forward(f, w(HEAD_KEYS[0]), w(HEAD_KEYS[1]), w(HEAD_KEYS[2]), w(HEAD_KEYS[3]), kc)
        ordl = oheads.ordinal_forward(f, w(HEAD_KEYS[4]), w(HEAD_KEYS[5]), w(HEAD_KEYS[6]), w(HEAD_KEYS[7]), ko)
        mu, lv = oheads.uncertainty_forward(f, *[w(k) for k in HEAD_KEYS[8:14]], ku)
        kan = okan.severity_forward(f, [tuple(w(k) for k in HEAD_KEYS[14 + 3 * l:17 + 3 * l]) for l in range(3)], sdg['kan_module.kan_layers.0.knots'])
        outs = {'cls': cls, 'ord': ordl, 'mu': mu, 'lv': lv, 'kan': kan}
        torch.autograd.backward([outs[k] for k in ups], [ups[k] for k in ups])
        return outs, f.grad, [sdg[k].grad for k in HEAD_KEYS]

    def ours(drop_p, seed=0, offset=0):
        params, knots, ws, f, out = _run_tail_forward(tail_lib, sd, feat, train=True, drop_p=drop_p, seed=seed, offset=offset)
        grads = [np.full_like(p, np.nan) for p in params]
        gtable = (ctypes.c_void_p * 23)(*[g.ctypes.data for g in grads])
        dfeat, dws = np.full_like(f, np.nan), np.full_like(ws, np.nan)
        u = {k: np.ascontiguousarray(v.numpy(), dtype=F) for k, v in ups.items()}
        tail_lib.tail_backward(vp(knots), vp(f), vp(ws), batch, drop_p, vp(out['h']), vp(out['a1']), vp(out['a2']), vp(out['lv']), vp(out['kan']),
                               vp(u['cls']), vp(u['ord']), vp(u['mu']), vp(u['lv']), vp(u['kan']), vp(dfeat), vp(dws), gtable)
        return out, dfeat, grads

    def compare(out, dfeat, grads, ref_out, ref_df, ref_grads):
        for k in ups:
            want = ref_out[k].detach().numpy()
            assert np.abs(out[k] - want).max() <= 2e-5 * max(1.0, float(np.abs(want).max())), k
        assert np.abs(dfeat - ref_df.numpy()).max() <= 3e-5 * float(ref_df.abs().max())
        for key, got, want in zip(HEAD_KEYS, grads, ref_grads):
            want = want.numpy()
            assert np.isfinite(got).all(), key
            assert np.abs(got - want).max() <= 3e-5 * float(np.abs(want).max()) + 1e-7, (key, float(np.abs(got - want).max()), float(np.abs(want).max()))

    out, dfeat, grads = ours(0.0)
    compare(out, dfeat, grads, *oracle())
    clamped = np.abs(out['lv'][:, 0]) == 10.0
    assert clamped.any() and not clamped.all()
    # ---- Dropout p = 0.3: recover the mask from the saved hidden activations (zero <=> dropped or ReLU-gated)
    out, dfeat, grads = ours(0.3, seed=1234, offset=77)
    with torch.no_grad():
        hidden = torch.cat([oheads.mlp_hidden(feat, sd[HEAD_KEYS[i]], sd[HEAD_KEYS[i + 1]]) for i in (0, 4, 8)], dim=1).numpy()
    live = hidden > 1e-4
    ratio = out['h'][live] / hidden[live]
    kept = ratio > 0.5
    assert np.abs(ratio[kept] - 1.0 / 0.7).max() <= 1e-4 and not ratio[~kept].any()
    assert abs((~kept).mean() - 0.3) < 0.04, (~kept).mean()
    keep = np.where(live, np.where(out['h'] != 0, 1.0 / 0.7, 0.0), 1.0 / 0.7).astype(F)
    keep3 = tuple(torch.from_numpy(np.ascontiguousarray(keep[:, i * 128:(i + 1) * 128])) for i in range(3))
    compare(out, dfeat, grads, *oracle(keep3))
    out2 = _run_tail_forward(tail_lib, sd, feat, train=True, drop_p=0.3, seed=1234, offset=78)[-1]
    assert not np.array_equal(out2['h'] != 0, out['h'] != 0)            # another offset, another mask


# ------------------------------------------------------------------------------------------ attention probabilities, loss reduction, gradient norm
def test_attention_probabilities_kernel_on_the_host(tmp_path):
    """`attn_probs_kernel` (csrc/attention.cu: softmax(q k^T / 8) per (image, head, query) as an explicit tensor for the reference's
    attention rollout, explainability/attention_maps.py:46-80; warp per query row, two shuffle reductions) against torch on the same
    bf16 qkv tensor in timm's `reshape(B, N, 3, H, d)` layout."""
    a = read('attention.cu')
    body = ('namespace {\n' + between(a, 'constexpr int kTok = 197', '}  // namespace') + '}\n' + r'''
extern "C" void probs(const uint16_t* qkv, float* out, int batch) {
  const long long warps = static_cast<long long>(batch) * kHeads * kTok;
  EmuDim g; g.x = static_cast<unsigned>((warps * 32 + 255) / 256); EmuDim b; b.x = 256;
  emu_launch(g, b, 0, [=] { attn_probs_kernel(reinterpret_cast<const __nv_bfloat16*>(qkv), out, batch); });
}
''')
    lib = compile_host(tmp_path, 'attn', body)
    lib.probs.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    torch.manual_seed(0)
    qkv = (torch.randn(1, 197, 576) * 1.5).to(torch.bfloat16)
    out = np.full((1, 3, 197, 197), np.nan, F)
    lib.probs(vp(np.ascontiguousarray(qkv.view(torch.int16).numpy())), vp(out), 1)
    t = qkv.float().reshape(1, 197, 3, 3, 64).permute(2, 0, 3, 1, 4)
    want = torch.softmax(t[0] @ t[1].transpose(-1, -2) * 0.125, dim=-1).numpy()
    assert np.abs(out - want).max() <= 2e-6 and np.abs(out.sum(-1) - 1.0).max() <= 1e-5


def test_joint_loss_kernel_with_its_block_reduction_on_the_host(tmp_path):
    """The whole `joint_loss_kernel` (per-sample terms -> warp sums -> shared partials -> one atomicAdd per block and term) and
    `loss_finalize_kernel` at a batch that spans three blocks, against the oracle (pinned to the reference's JointLoss)."""
    import sys
    sys.path.insert(0, ROOT)
    from oracle import losses as olosses
    h, c = read('heads.cu'), read('common.cuh')
    body = ('namespace {\n' + between(c, '__device__ __forceinline__ float warp_sum', '__device__ __forceinline__ float warp_max')
            + between(h, 'struct LossParams {', '// RoViTKAN.predict epilogue')
            + between(h, '// out = {cls, ord, unc, kan, cls + l_ord', '// dst[i] = src[i] * (g[term]') + '}\n' + r'''
extern "C" void loss(const float* cls, const float* ordl, const float* mu, const float* lv, const float* kan, const long long* yc,
                     const float* ys, const float* alpha, float gamma, int batch, float* sums, float* out5, float* d_cls, float* d_ord,
                     float* d_mu, float* d_lv, float* d_kan) {
  LossParams p;
  p.cls_logits = cls; p.num_classes = 4; p.ord_logits = ordl; p.mu = mu; p.log_var = lv; p.kan = kan; p.class_t = yc; p.sev_t = ys;
  p.alpha = alpha; p.gamma = gamma; p.batch = batch; p.sums = sums;
  p.d_cls = d_cls; p.d_ord = d_ord; p.d_mu = d_mu; p.d_lv = d_lv; p.d_kan = d_kan;
  for (int i = 0; i < 4; ++i) sums[i] = 0.0f;
  EmuDim g; g.x = (batch + 255) / 256; EmuDim b; b.x = 256;
  emu_launch(g, b, 0, [=] { joint_loss_kernel(p); });
  EmuDim one; EmuDim b32; b32.x = 32;
  emu_launch(one, b32, 0, [=] { loss_finalize_kernel(sums, 1.0f, 0.5f, 0.5f, out5); });
}
''')
    lib = compile_host(tmp_path, 'loss', body)
    P = ctypes.c_void_p
    lib.loss.argtypes = [P, P, P, P, P, P, P, P, ctypes.c_float, ctypes.c_int] + [P] * 7
    g = torch.Generator().manual_seed(1)
    B = 600
    o = {'cls_logits': torch.randn(B, 4, generator=g) * 2, 'ordinal_logits': torch.randn(B, 3, generator=g) * 2, 'mu': torch.randn(B, 1, generator=g),
         'log_var': torch.randn(B, 1, generator=g), 'kan_severity': torch.rand(B, 1, generator=g) * 3}
    y = torch.randint(0, 4, (B,), generator=g)
    alpha = torch.rand(4, generator=g) + 0.5
    og = {k: v.clone().requires_grad_(True) for k, v in o.items()}
    ref = olosses.joint(og, y, y, 4, alpha=alpha)
    ref['total_loss'].backward()
    n = lambda t, dt=F: np.ascontiguousarray(t.numpy(), dtype=dt)
    sums, out5 = np.zeros(4, F), np.zeros(5, F)
    d = [np.full((B, k), np.nan, F) for k in (4, 3, 1, 1, 1)]
    lib.loss(vp(n(o['cls_logits'])), vp(n(o['ordinal_logits'])), vp(n(o['mu'])), vp(n(o['log_var'])), vp(n(o['kan_severity'])), vp(n(y, np.int64)),
             vp(n(y.float())), vp(n(alpha)), 2.0, B, vp(sums), vp(out5), *[vp(a) for a in d])
    for i, k in enumerate(('cls_loss', 'ord_loss', 'unc_loss', 'kan_loss', 'total_loss')):
        assert abs(float(out5[i]) - float(ref[k].detach())) <= 3e-6 * max(1.0, abs(float(ref[k].detach()))), k
    for got, w, k in zip(d, (1.0, 1.0, 0.5, 0.5, 0.5), o):
        assert np.abs(w * got - og[k].grad.numpy()).max() <= 1e-5 * float(og[k].grad.abs().max()) + 1e-9, k


def test_gradient_norm_kernel_on_the_host(tmp_path):
    """`optim_norm_kernel` (csrc/optimizer.cu: sum of squares of every gradient over a padded chunk space, float4 loads where the
    pointer allows, warp + shared reduction, one atomicAdd per block) = what clip_grad_norm_ computes; a tensor without gradient and
    an unaligned gradient pointer are part of the table."""
    o = read('optimizer.cu')
    body = ('namespace {\n' + between(o, 'constexpr int kChunk', '// state: [0] running sum of squares') + '}\n' + r'''
extern "C" float sumsq(int n, const float** grads, const long long* numel) {
  OptTable T{};
  T.n = n;
  int chunks = 0;
  for (int i = 0; i < n; ++i) { T.g[i] = grads[i]; T.numel[i] = static_cast<int>(numel[i]); T.chunk_start[i] = chunks;
                                chunks += static_cast<int>((numel[i] + kChunk - 1) / kChunk); }
  T.chunk_start[n] = chunks;
  static float state[4];
  state[0] = 0.0f;
  EmuDim g; g.x = chunks; EmuDim b; b.x = kOptThreads;
  float* st = state;
  emu_launch(g, b, 0, [=] { optim_norm_kernel(T, st); });
  return state[0];
}
''')
    lib = compile_host(tmp_path, 'norm', body)
    lib.sumsq.restype = ctypes.c_float
    lib.sumsq.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    rng = np.random.default_rng(0)
    sizes = [5000, 17, 4096, 8193, 1]
    bufs = [rng.normal(0, 1, s + 1).astype(F) for s in sizes]
    grads = [b[:-1] for b in bufs]
    grads[3] = bufs[3][1:]                                              # 4-byte aligned only: the scalar path
    ptrs = [g.ctypes.data for g in grads]
    ptrs[1] = None                                                     # a parameter without gradient
    table = (ctypes.c_void_p * 5)(*ptrs)
    got = lib.sumsq(5, table, vp(np.array(sizes, np.int64)))
    want = sum(float((g.astype(np.float64) ** 2).sum()) for i, g in enumerate(grads) if i != 1)
    assert abs(got - want) <= 2e-6 * want


# ------------------------------------------------------------------------------------------ trunk plumbing kernels
def test_weight_shadow_token_and_column_sum_kernels_on_the_host(tmp_path):
    """The HBM-bound helpers around the tcgen05 trunk kernels, emulated: `cast_multi_kernel` (all bf16 / transposed-bf16 / fp16
    weight shadows in one launch, tile index by binary search over a job table), `token_table_kernel` (cls_token + pos_embed and
    patch bias + pos_embed folded into one additive table), `token_grad_reduce_kernel` (their gradients) and `colsum_kernel`
    (bias gradients from fp32 / bf16 rows)."""
    k, c = read('encoder_kernels.cu'), read('common.cuh')
    kh = read('kernels.h')
    body = (between(kh, 'struct RvkCastJob', 'int rvk_cast_multi_launch') + 'namespace {\n'
            + between(c, '__device__ __forceinline__ float warp_sum', '__device__ __forceinline__ float warp_max')
            + between(k, 'constexpr int kD = 192;', '// ------------------------------------------------------------------ patch extraction')
            + between(k, '// table[0] = cls_token + pos[0]', '__global__ void cast_bf16_kernel')
            + between(k, '// ------------------------------------------------------------------ column sums', '}  // namespace') + '}\n' + r'''
extern "C" void cast_multi(int n, const float** src, void** dst, const int* rows, const int* cols, const int* mode) {
  RvkCastTable T{};
  T.n = n;
  int tiles = 0;
  for (int i = 0; i < n; ++i) {          // rvk_cast_multi_launch
    T.job[i].src = src[i]; T.job[i].dst = dst[i]; T.job[i].rows = rows[i]; T.job[i].cols = cols[i]; T.job[i].mode = mode[i];
    T.job[i].tile_start = tiles;
    tiles += ((rows[i] + 31) / 32) * ((cols[i] + 31) / 32);
  }
  EmuDim g; g.x = tiles; EmuDim b; b.x = 32; b.y = 8;
  emu_launch(g, b, 0, [=] { cast_multi_kernel(T); });
}
extern "C" void token_table(const float* cls, const float* pos, const float* pbias, float* table) {
  EmuDim g; g.x = (kTok * kD + 255) / 256; EmuDim b; b.x = 256;
  emu_launch(g, b, 0, [=] { token_table_kernel(cls, pos, pbias, table); });
}
extern "C" void token_grads(const float* dx0, int batch, int b_per_block, float* dpos, float* dcls, float* dpbias) {
  EmuDim g; g.x = kTok; g.y = (batch + b_per_block - 1) / b_per_block; EmuDim b; b.x = kD;
  emu_launch(g, b, 0, [=] { token_grad_reduce_kernel(dx0, batch, b_per_block, dpos, dcls, dpbias); });
}
extern "C" void colsum(int bf16, const void* src, long long ld, int rows, int cols, float* out, float scale, int rows_per_block) {
  EmuDim g; g.x = (rows + rows_per_block - 1) / rows_per_block; EmuDim b; b.x = 256;
  if (bf16) emu_launch(g, b, 0, [=] { colsum_kernel<true>(src, ld, rows, cols, out, scale, rows_per_block); });
  else emu_launch(g, b, 0, [=] { colsum_kernel<false>(src, ld, rows, cols, out, scale, rows_per_block); });
}
''')
    lib = compile_host(tmp_path, 'plumb', body)
    P, I = ctypes.c_void_p, ctypes.c_int
    lib.cast_multi.argtypes = [I, P, P, P, P, P]
    lib.token_table.argtypes = [P, P, P, P]
    lib.token_grads.argtypes = [P, I, I, P, P, P]
    lib.colsum.argtypes = [I, P, ctypes.c_longlong, I, I, P, ctypes.c_float, I]
    rng = np.random.default_rng(0)
    # ---- weight shadows: qkv-like [576,192] -> bf16, fc2-like [192,768] -> fp16, proj-like [192,192] -> transposed bf16, ragged [70,45] x 3
    shapes = [(576, 192, 0), (192, 768, 2), (192, 192, 1), (70, 45, 0), (70, 45, 1), (70, 45, 2)]
    srcs = [rng.normal(0, 0.05, (r, c)).astype(F) for r, c, _ in shapes]
    srcs[3][0, 0], srcs[5][0, 0], srcs[5][0, 1] = 1e-40, 3e-6, 70000.0                 # denormal input, fp16 subnormal, fp16 overflow
    dsts = [np.full((c, r) if m == 1 else (r, c), 0xffff, np.uint16) for r, c, m in shapes]
    arr = lambda vals, dt: np.array(vals, dt)
    lib.cast_multi(len(shapes), (ctypes.c_void_p * 6)(*[s.ctypes.data for s in srcs]), (ctypes.c_void_p * 6)(*[d.ctypes.data for d in dsts]),
                   vp(arr([s[0] for s in shapes], np.int32)), vp(arr([s[1] for s in shapes], np.int32)), vp(arr([s[2] for s in shapes], np.int32)))
    for (r, c, m), s, d in zip(shapes, srcs, dsts):
        t = torch.from_numpy(s)
        if m == 2:
            assert np.array_equal(d, t.to(torch.float16).view(torch.int16).numpy().astype(np.uint16)), (r, c, m)
        else:
            want = (t.t().contiguous() if m == 1 else t).to(torch.bfloat16).view(torch.int16).numpy().astype(np.uint16)
            assert np.array_equal(d, want), (r, c, m)
    # ---- additive token table and its gradients
    cls, pos, pb = (rng.normal(0, 1, s).astype(F) for s in ((192,), (197, 192), (192,)))
    table = np.full((197, 192), np.nan, F)
    lib.token_table(vp(cls), vp(pos), vp(pb), vp(table))
    want = pos.copy(); want[0] += cls; want[1:] += pb
    assert np.array_equal(table, want)
    dx0 = rng.normal(0, 1, (5, 197, 192)).astype(F)
    dpos, dcls, dpb = np.zeros((197, 192), F), np.zeros(192, F), np.zeros(192, F)
    lib.token_grads(vp(dx0), 5, 2, vp(dpos), vp(dcls), vp(dpb))
    assert np.abs(dpos - dx0.sum(0)).max() <= 1e-5 and np.abs(dcls - dx0[:, 0].sum(0)).max() <= 1e-5
    assert np.abs(dpb - dx0[:, 1:].sum((0, 1))).max() <= 2e-5 * np.abs(dx0[:, 1:].sum((0, 1))).max()
    # ---- column sums (bias gradients): fp32 rows of 192, bf16 rows of 768 and 576, accumulate into `out` with a scale
    x32 = rng.normal(0, 1, (333, 192)).astype(F)
    out = np.ones(192, F)
    lib.colsum(0, vp(x32), 192, 333, 192, vp(out), 0.5, 64)
    assert np.abs(out - (1 + 0.5 * x32.sum(0))).max() <= 2e-5 * np.abs(x32.sum(0)).max()
    for cols in (768, 576):
        xb = torch.from_numpy(rng.normal(0, 1, (200, cols)).astype(F)).to(torch.bfloat16)
        out = np.zeros(cols, F)
        lib.colsum(1, vp(np.ascontiguousarray(xb.view(torch.int16).numpy())), cols, 200, cols, vp(out), 1.0, 50)
        assert np.abs(out - xb.float().sum(0).numpy()).max() <= 2e-5 * float(xb.float().sum(0).abs().max())


# ------------------------------------------------------------------------------------------ the per-layer heads path
def test_per_layer_linear_kernels_on_the_host(tmp_path):
    """`rvk_linear_forward` / `rvk_linear_backward` as launched for a head layer outside the fused tail (a module called on its own,
    e.g. OrdinalHead.predict_probabilities heads.py:45-67, or a non-default `kan_layers` model): `sgemm_kernel<1>` with the fused
    bias / ReLU / clamp epilogue, `epilogue_grad_kernel` (mask recovered from the OUTPUT), dx = g W, the split-K weight gradient with
    atomics and `colsum_small_kernel` -- against torch autograd, ragged sizes."""
    h, c = read('heads.cu'), read('common.cuh')
    body = ('namespace {\n' + between(c, '__device__ __forceinline__ uint4 philox4x32_10', '__device__ __forceinline__ uint32_t smem_u32')
            + between(c, '__device__ __forceinline__ float warp_sum', '__device__ __forceinline__ float warp_max')
            + between(h, 'struct SgemmParams {', '// ------------------------------------------------------------------ joint loss') + '}\n' + r'''
static void sgemm(const float* A, long long lda, int transA, const float* B, long long ldb, int transB, float* C, long long ldc, int M, int N,
                  int K, const float* bias, int relu, float lo, float hi, int accumulate, int split_k) {
  SgemmParams p{};
  p.A = A; p.lda = lda; p.transA = transA; p.B = B; p.ldb = ldb; p.transB = transB; p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.K = K;
  p.bias = bias; p.relu = relu; p.clamp_lo = lo; p.clamp_hi = hi; p.accumulate = accumulate;
  int splits = 1;
  if (split_k) {                                             // rvk_sgemm_launch
    const int tiles = ((M + 63) / 64) * ((N + 63) / 64);
    splits = (148 + tiles - 1) / tiles;
    const int max_splits = (K + 63) / 64;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
  }
  p.k_per_split = ((K + splits - 1) / splits + 15) / 16 * 16;
  splits = (K + p.k_per_split - 1) / p.k_per_split;
  EmuDim g; g.x = (M + 15) / 16; g.y = (N + 63) / 64; g.z = splits; EmuDim b; b.x = 256;
  emu_launch(g, b, 0, [=] { sgemm_kernel<1>(p); });
}
extern "C" void linear_fwd(const float* x, const float* w, const float* bias, int batch, int n_in, int n_out, int relu, float lo, float hi, float* y) {
  sgemm(x, n_in, 0, w, n_in, 1, y, n_out, batch, n_out, n_in, bias, relu, lo, hi, 0, 0);
}
extern "C" void linear_bwd(const float* x, const float* w, const float* y, const float* gy, int batch, int n_in, int n_out, int relu, float lo,
                           float hi, float* dx, float* dw, float* db, float* gpre_ws) {
  const float* gpre = gy;
  if (relu || lo < hi) {                                     // rvk_linear_backward
    EmuDim g; g.x = (batch * n_out + 255) / 256; EmuDim b; b.x = 256;
    emu_launch(g, b, 0, [=] { epilogue_grad_kernel(y, gy, gpre_ws, relu, 1.0f, lo, hi, batch * n_out); });
    gpre = gpre_ws;
  }
  sgemm(gpre, n_out, 0, w, n_in, 0, dx, n_in, batch, n_in, n_out, nullptr, 0, 0.f, 0.f, 0, 0);
  sgemm(gpre, n_out, 1, x, n_in, 0, dw, n_in, n_out, n_in, batch, nullptr, 0, 0.f, 0.f, 1, 1);
  EmuDim g; g.x = n_out; EmuDim b; b.x = 256;
  emu_launch(g, b, 0, [=] { colsum_small_kernel(gpre, n_out, batch, n_out, db); });
}
''')
    lib = compile_host(tmp_path, 'lin', body)
    P, I, Fl = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    lib.linear_fwd.argtypes = [P, P, P, I, I, I, I, Fl, Fl, P]
    lib.linear_bwd.argtypes = [P, P, P, P, I, I, I, I, Fl, Fl, P, P, P, P]
    torch.manual_seed(0)
    for batch, n_in, n_out, relu, clamp in ((37, 192, 128, 1, None), (37, 128, 4, 0, None), (150, 128, 1, 0, (-10.0, 10.0))):
        lin = torch.nn.Linear(n_in, n_out)
        if clamp:
            with torch.no_grad():
                lin.weight.mul_(40.0)
        x = torch.randn(batch, n_in, requires_grad=True)
        y_ref = lin(x)
        y_ref = torch.relu(y_ref) if relu else y_ref
        y_ref = torch.clamp(y_ref, *clamp) if clamp else y_ref
        gy = torch.randn(batch, n_out)
        y_ref.backward(gy)
        lo, hi = clamp if clamp else (0.0, 0.0)
        n = lambda t: np.ascontiguousarray(t.detach().numpy(), dtype=F)
        y = np.full((batch, n_out), np.nan, F)
        lib.linear_fwd(vp(n(x)), vp(n(lin.weight)), vp(n(lin.bias)), batch, n_in, n_out, relu, lo, hi, vp(y))
        assert np.abs(y - n(y_ref)).max() <= 2e-5 * float(y_ref.detach().abs().max())
        dx, ws = np.full((batch, n_in), np.nan, F), np.full((batch, n_out), np.nan, F)
        dw, db = np.zeros((n_out, n_in), F), np.zeros(n_out, F)
        lib.linear_bwd(vp(n(x)), vp(n(lin.weight)), vp(y), vp(n(gy)), batch, n_in, n_out, relu, lo, hi, vp(dx), vp(dw), vp(db), vp(ws))
        for got, want in ((dx, x.grad), (dw, lin.weight.grad), (db, lin.bias.grad)):
            assert np.abs(got - n(want)).max() <= 3e-5 * float(want.abs().max()) + 1e-7
        if clamp:
            sat = np.abs(y[:, 0]) == 10.0
            assert sat.any() and not sat.all() and not dx[sat].any()             # clamp: zero gradient outside (-10, 10)


def test_tensor_core_kan_weight_split_kernels_on_the_host(tmp_path):
    """The weight operands of the tcgen05 KAN kernels: `kan_split_weights_kernel` ([64 outputs][in * 8] for the forward / dW
    kernels) and `kan_split_weights_rows_kernel` ([in * 8][64] for the dx kernel) pack W[i,o,k] (k < 7) and Wl[o,i] (k = 7) like the
    fp32 path and split every value into bf16 hi + lo: hi + lo must reproduce the fp32 weight to 2^-16 relative (three MMAs per product
    then give an fp32-grade result) and the padding must be exact zeros."""
    t = read('kan_tc.cuh')
    body = ('namespace {\n' + between(t, '// WpT (fp32 [out_pad=64][kp]) -> bf16 hi / lo', '// The 8 packed activations')
            + between(t, '// Wp (fp32 [kp][out_pad=64]) -> bf16 hi / lo', '__global__ void __launch_bounds__(kTcThreads, 1)') + '}\n' + r'''
extern "C" void split(int rows_major, const float* spline, const float* lin_w, int n_in, int n_out, int kp, uint16_t* hi, uint16_t* lo) {
  EmuDim g; g.x = 7; EmuDim b; b.x = 256;
  auto* h = reinterpret_cast<__nv_bfloat16*>(hi); auto* l = reinterpret_cast<__nv_bfloat16*>(lo);
  if (rows_major) emu_launch(g, b, 0, [=] { kan_split_weights_rows_kernel(spline, lin_w, n_in, n_out, kp, h, l); });
  else emu_launch(g, b, 0, [=] { kan_split_weights_kernel(spline, lin_w, n_in, n_out, kp, h, l); });
}
''')
    lib = compile_host(tmp_path, 'split', body)
    P, I = ctypes.c_void_p, ctypes.c_int
    lib.split.argtypes = [I, P, P, I, I, I, P, P]
    rng = np.random.default_rng(0)
    n_in, n_out = 72, 50                                    # padded to 80 inputs (kp = 640) and 64 outputs
    kp = 80 * 8
    spline, lin_w = rng.normal(0, 0.1, (n_in, n_out, 7)).astype(F), rng.normal(0, 0.1, (n_out, n_in)).astype(F)
    want = np.zeros((64, kp), F)                             # [o][i * 8 + k]
    for k in range(7):
        want[:n_out, np.arange(n_in) * 8 + k] = spline[:, :, k].T
    want[:n_out, np.arange(n_in) * 8 + 7] = lin_w
    f32 = lambda a: (a.astype(np.uint32) << 16).view(F)
    for rows_major in (0, 1):
        hi, lo = np.full(64 * kp, 0xffff, np.uint16), np.full(64 * kp, 0xffff, np.uint16)
        lib.split(rows_major, vp(spline), vp(lin_w), n_in, n_out, kp, vp(hi), vp(lo))
        w = want.T if rows_major else want
        got = (f32(hi).astype(np.float64) + f32(lo).astype(np.float64)).reshape(w.shape)
        assert np.abs(got - w).max() <= 2.0 ** -16 * np.abs(w).max()
        assert not hi.reshape(w.shape)[w == 0].any() and not lo.reshape(w.shape)[w == 0].any()
        assert np.array_equal(hi.reshape(w.shape), torch.from_numpy(np.ascontiguousarray(w)).to(torch.bfloat16).view(torch.int16).numpy().astype(np.uint16))


# ------------------------------------------------------------------------------------------ a tcgen05 kernel, functionally
def test_weight_gradient_tcgen05_kernel_on_the_host(tmp_path):
    """`gemm_tn_kernel<192, 4>` (csrc/gemm_tn.cuh: dW = A^T B on tcgen05 / TMEM, both operands MN-major through TMA with the
    128-byte swizzle, split over the rows, fused bias gradient through a constant ones panel, red.global.add epilogue) run under a
    FUNCTIONAL emulation of the Blackwell pieces it uses (tests/host_emu/tcgen05_host_emu.h: mbarriers with transaction counts,
    swizzled TMA tiles, tensor memory, tcgen05.mma decoded from the shared-memory / instruction descriptors that common.cuh's own,
    verbatim builders produce).  Warp roles, pipeline phases, descriptor strides, the ones-panel column and the split-K epilogue all
    execute; the row count is ragged (zero-filled TMA rows) and the result accumulates into a non-zero C."""
    c, t = read('common.cuh'), read('gemm_tn.cuh')
    kernel = between(t, 'struct GemmTnParams {', '#endif  // __CUDACC__').replace('#ifdef __CUDACC__', '')
    dyn = 'extern __shared__ uint8_t smem_raw[];'
    assert kernel.count(dyn) == 1
    kernel = kernel.replace(dyn, 'uint8_t* smem_raw = static_cast<uint8_t*>(emu_dynamic_smem());')
    body = ('#include "tcgen05_host_emu.h"\nnamespace {\n'
            + between(c, '// Shared-memory matrix descriptor (64-bit).', '// byte offset of 16-byte chunk')      # umma_smem_desc + umma_idesc_bf16, verbatim
            + kernel + '}\n' + r'''
extern "C" void gemm_tn(const uint16_t* A, long long lda, const uint16_t* B, long long ldb, float* C, long long ldc, int M, int P, int Q,
                        float scale, float* colsum, int splits_wanted) {
  constexpr int kBQ = 192, kStages = 4;
  using L = GemmTnSmem<kBQ, kStages>;
  const CUtensorMap tmA = emu_make_tmap_2d(A, 2, M, P, lda, 64, 64), tmB = emu_make_tmap_2d(B, 2, M, Q, ldb, 64, 64);   // rvk_gemm_tn_launch
  const int tiles = ((P + 127) / 128) * ((Q + kBQ - 1) / kBQ);
  const int total_chunks = (M + 63) / 64;
  int splits = splits_wanted;
  if (splits > total_chunks) splits = total_chunks;
  if (splits < 1) splits = 1;
  GemmTnParams p;
  p.M = M; p.P = P; p.Q = Q; p.ldc = static_cast<int>(ldc);
  p.chunks_per_split = (total_chunks + splits - 1) / splits;
  p.C = C; p.scale = scale; p.colsum = colsum;
  splits = (total_chunks + p.chunks_per_split - 1) / p.chunks_per_split;
  EmuDim g; g.x = tiles; g.y = splits; EmuDim b; b.x = kTnThreads;
  for (int i = 0; i < 16; ++i) emu_named_n[i] = 0;
  emu_launch(g, b, L::kTotal, [=] { gemm_tn_kernel<kBQ, kStages>(tmA, tmB, p); });
}
''')
    lib = compile_host(tmp_path, 'gemm_tn', body)
    P_, I, L_, Fl = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float
    lib.gemm_tn.argtypes = [P_, L_, P_, L_, P_, L_, I, I, I, Fl, P_, I]
    g = torch.Generator().manual_seed(0)
    M, P, Q = 330, 192, 192                                  # 330 rows: the sixth 64-row chunk is mostly zero-filled by TMA
    A = (torch.randn(M, P, generator=g)).to(torch.bfloat16)
    B = (torch.randn(M, Q, generator=g)).to(torch.bfloat16)
    bits = lambda x: np.ascontiguousarray(x.view(torch.int16).numpy())
    want = A.double().t() @ B.double()
    for splits, with_colsum in ((1, False), (3, True), (6, True)):
        C0 = torch.randn(P, Q, generator=g)
        C = C0.numpy().copy()
        cs = np.full(P, 2.0, F)
        lib.gemm_tn(vp(bits(A)), P, vp(bits(B)), Q, vp(C), Q, M, P, Q, 0.5, vp(cs) if with_colsum else None, splits)
        ref = C0.double() + 0.5 * want
        assert np.abs(C - ref.numpy()).max() <= 1e-5 * float(ref.abs().max()), (splits, float(np.abs(C - ref.numpy()).max()))
        if with_colsum:
            assert np.abs(cs - (2.0 + 0.5 * A.double().sum(0)).numpy()).max() <= 1e-5 * float(A.double().sum(0).abs().max() + 2)


@pytest.fixture(scope='module')
def gemm_nt_lib(tmp_path_factory):
    c, t = read('common.cuh'), read('gemm_nt.cuh')
    kernel = between(t, 'enum GemmEpilogue', '#endif  // __CUDACC__').replace('#ifdef __CUDACC__', '')
    dyn = 'extern __shared__ uint8_t smem_raw[];'
    assert kernel.count(dyn) == 1
    kernel = kernel.replace(dyn, 'uint8_t* smem_raw = static_cast<uint8_t*>(emu_dynamic_smem());')
    gelu = between(c, '// Phi(-a) for a >= 0', '// ----------------------------------------------------------------------------- mbarrier')
    body = ('#include "tcgen05_host_emu.h"\nnamespace {\n'
            + between(c, '// Shared-memory matrix descriptor (64-bit).', '// ----------------------------------------------------------------------------- CTA pairs')   # descriptor builders + sw128_offset, verbatim
            + between(c, 'constexpr uint32_t kUmmaDescHiSw128', 'template <int G>\n__device__ __forceinline__ void umma_f16_split')                                   # umma_desc_lo, verbatim
            + 'template <int G> inline void umma_f16_split(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) {\n'
              '  static_assert(G == 1, "single-CTA emulation"); emu_umma_split(kUmmaDescHiSw128, d, a_lo, b_lo, idesc, acc); }\n'
            + gelu + between(c, '__host__ __device__ __forceinline__ size_t xt_offset', '#endif  // __CUDACC__')
            + kernel + '}\n' + r'''
// tensor maps, grid and stage choice of launch_nt_stages / launch_nt (gemm.cu); `grid_cap` stands for the SM count
template <int MODE, int STAGES>
static void run_nt(const uint16_t* A, const uint16_t* B, void* out, void* out2, const void* aux, GemmNtParams p, int grid_cap) {
  constexpr int kBN = 192;
  using L = GemmNtSmem<kBN, STAGES>;
  const bool out_f32 = (MODE == EPI_F32 || MODE == EPI_RES_LN);
  const CUtensorMap tmA = emu_make_tmap_2d(A, 2, p.M, p.K, p.K, 128, 64), tmB = emu_make_tmap_2d(B, 2, p.N, p.K, p.K, kBN, 64);
  CUtensorMap tmOut = (MODE == EPI_RES_LN && p.out_tiled != nullptr) ? tmA : emu_make_tmap_2d(out, out_f32 ? 4 : 2, p.M, p.N, p.N, 32, out_f32 ? 32 : 64);
  CUtensorMap tmOut2 = tmOut, tmAux = tmOut;
  if (p.has_out2) tmOut2 = emu_make_tmap_2d(out2, 2, p.M, p.N, p.N, 32, 64);
  if (MODE == EPI_DGELU) tmAux = emu_make_tmap_2d(aux, 2, p.M, p.N, p.N, 128, 64);
  else if (MODE == EPI_RES_LN && p.has_res && p.res_table == nullptr && p.out_tiled == nullptr) tmAux = emu_make_tmap_2d(aux, 4, p.M, p.N, p.N, 128, 32);
  const int tiles = ((p.M + 127) / 128) * (p.N / kBN);
  EmuDim g; g.x = tiles < grid_cap ? tiles : grid_cap; EmuDim b; b.x = kGemmThreads;
  for (int i = 0; i < 16; ++i) emu_named_n[i] = 0;
  emu_launch(g, b, L::kTotal, [=] { gemm_nt_kernel<kBN, MODE, STAGES>(tmA, tmB, tmOut, tmOut2, tmAux, p); });
}
extern "C" void gemm_nt(int mode, const uint16_t* A, const uint16_t* B, void* out, void* out2, const void* aux, int M, int N, int K,
                        const float* bias, const float* gamma, const float* beta, const float* res_table, int table_rows, float eps,
                        float* mean, float* rstd, int has_out2, int has_res, int grid_cap) {
  GemmNtParams p{};
  p.M = M; p.N = N; p.K = K; p.bias = bias; p.gamma = gamma; p.beta = beta; p.res_table = res_table; p.table_rows = table_rows;
  p.ln_eps = eps; p.mean_out = mean; p.rstd_out = rstd; p.has_out2 = has_out2; p.has_res = has_res;
  const bool two = K <= 192 && mode == EPI_DGELU;
  if (mode == EPI_BF16) run_nt<EPI_BF16, 3>(A, B, out, out2, aux, p, grid_cap);
  else if (mode == EPI_GELU) run_nt<EPI_GELU, 3>(A, B, out, out2, aux, p, grid_cap);
  else if (mode == EPI_DGELU) { if (two) run_nt<EPI_DGELU, 2>(A, B, out, out2, aux, p, grid_cap); else run_nt<EPI_DGELU, 3>(A, B, out, out2, aux, p, grid_cap); }
  else if (mode == EPI_F32) run_nt<EPI_F32, 3>(A, B, out, out2, aux, p, grid_cap);
  else run_nt<EPI_RES_LN, 3>(A, B, out, out2, aux, p, grid_cap);
}
// inference: the fp32 token stream leaves (and a residual arrives) in the tiled layout of common.cuh, straight from / to registers
extern "C" void gemm_nt_tiled(const uint16_t* A, const uint16_t* B, void* ln_out, int M, int K, const float* bias, const float* gamma,
                              const float* beta, const float* res_table, int table_rows, const float* res_tiled, float* out_tiled, float eps,
                              int grid_cap) {
  GemmNtParams p{};
  p.M = M; p.N = 192; p.K = K; p.bias = bias; p.gamma = gamma; p.beta = beta; p.res_table = res_table; p.table_rows = table_rows;
  p.ln_eps = eps; p.has_out2 = 1; p.has_res = 1; p.res_tiled = res_tiled; p.out_tiled = out_tiled;
  run_nt<EPI_RES_LN, 3>(A, B, nullptr, ln_out, nullptr, p, grid_cap);
}
''')
    lib = compile_host(tmp_path_factory.mktemp('gemm_nt'), 'gemm_nt', body)
    P_, I, Fl = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    lib.gemm_nt.argtypes = [I, P_, P_, P_, P_, P_, I, I, I, P_, P_, P_, P_, I, Fl, P_, P_, I, I, I]
    lib.gemm_nt_tiled.argtypes = [P_, P_, P_, I, I, P_, P_, P_, P_, I, P_, P_, Fl, I]
    return lib


def _bf16_bits(x):
    return np.ascontiguousarray(x.view(torch.int16).numpy())


def _from_bits(a):
    return torch.from_numpy(a.astype(np.int16)).view(torch.bfloat16).float()


def test_forward_gemm_tcgen05_kernel_bf16_epilogue_on_the_host(gemm_nt_lib):
    """`gemm_nt_kernel<192, EPI_BF16, 3>` (the qkv projection: persistent, warp-specialised -- TMA producer, UMMA issuer over
    K-major swizzled operands through split descriptors, epilogue-panel producer, twelve epilogue warps in three teams, double-
    buffered TMEM accumulator, swizzled staging panels leaving by TMA store) under the functional emulation: M = 330 (ragged last
    tile, clipped stores), N = 576 = three N tiles, K = 192, on 2 "SMs" so that every CTA walks several tiles."""
    g = torch.Generator().manual_seed(0)
    M, N, K = 330, 576, 192
    A = torch.randn(M, K, generator=g).to(torch.bfloat16)
    W = (torch.randn(N, K, generator=g) * 0.1).to(torch.bfloat16)
    bias = torch.randn(N, generator=g)
    out = np.full((M, N), 0x7fc0, np.uint16)
    gemm_nt_lib.gemm_nt(0, vp(_bf16_bits(A)), vp(_bf16_bits(W)), vp(out), None, None, M, N, K, vp(bias.numpy()), None, None, None, 0, 0.0, None, None, 0, 0, 2)
    want = (A.double() @ W.double().t() + bias.double()).float()
    got = _from_bits(out)
    assert float((got - want).abs().max()) <= 2.0 ** -8 * float(want.abs().max())            # one bf16 rounding of the output
    assert float((got - want.to(torch.bfloat16).float()).abs().max()) <= 2.0 ** -7 * float(want.abs().max()) and torch.isfinite(got).all()


def test_forward_gemm_tcgen05_kernel_fused_epilogues_on_the_host(gemm_nt_lib):
    """The other epilogue modes of `gemm_nt_kernel` under the functional emulation, against torch on the same bf16 operands:
    EPI_GELU with both outputs (fc1: z = acc + bias and h = gelu(z), exact-erf GELU), EPI_DGELU (fc2's input gradient: acc * gelu'(z)
    with the saved z panels arriving by TMA into the staging slots, two operand stages / eight slots), EPI_F32, and EPI_RES_LN as the
    patch embedding runs it in training (acc + bias + token table by row % 197, fp32 stream out by TMA, LayerNorm1 of block 0 as a second
    bf16 output, per-row mean / rstd)."""
    g = torch.Generator().manual_seed(1)
    M = 300
    # ---- fc1: [M,192] x [768,192]^T, GELU, z kept
    A = torch.randn(M, 192, generator=g).to(torch.bfloat16)
    W = (torch.randn(768, 192, generator=g) * 0.1).to(torch.bfloat16)
    bias = torch.randn(768, generator=g) * 0.5
    h, z = np.full((M, 768), 0x7fc0, np.uint16), np.full((M, 768), 0x7fc0, np.uint16)
    gemm_nt_lib.gemm_nt(1, vp(_bf16_bits(A)), vp(_bf16_bits(W)), vp(h), vp(z), None, M, 768, 192, vp(bias.numpy()), None, None, None, 0, 0.0, None, None, 1, 0, 3)
    zw = (A.double() @ W.double().t() + bias.double()).float()
    assert float((_from_bits(z) - zw).abs().max()) <= 2.0 ** -8 * float(zw.abs().max())
    hw = torch.nn.functional.gelu(zw)
    assert float((_from_bits(h) - hw).abs().max()) <= 2.0 ** -8 * float(hw.abs().max()) + 1e-5
    # ---- fc2 dgrad: dz = (dy W2) * gelu'(z): A = dy [M,192], B = W2^T as [768,192], aux = z
    dy = torch.randn(M, 192, generator=g).to(torch.bfloat16)
    W2T = (torch.randn(768, 192, generator=g) * 0.1).to(torch.bfloat16)
    zsaved = _from_bits(z).to(torch.bfloat16)
    dz = np.full((M, 768), 0x7fc0, np.uint16)
    gemm_nt_lib.gemm_nt(2, vp(_bf16_bits(dy)), vp(_bf16_bits(W2T)), vp(dz), None, vp(_bf16_bits(zsaved)), M, 768, 192, None, None, None, None, 0, 0.0, None, None, 0, 0, 2)
    zz = zsaved.double().requires_grad_(True)
    torch.nn.functional.gelu(zz).sum().backward()
    want = ((dy.double() @ W2T.double().t()) * zz.grad).float()
    assert float((_from_bits(dz) - want).abs().max()) <= 2.0 ** -8 * float(want.abs().max()) + 1e-4
    # ---- fp32 output
    out32 = np.full((M, 192), np.nan, F)
    Wq = (torch.randn(192, 192, generator=g) * 0.1).to(torch.bfloat16)
    b192 = torch.randn(192, generator=g)
    gemm_nt_lib.gemm_nt(3, vp(_bf16_bits(A)), vp(_bf16_bits(Wq)), vp(out32), None, None, M, 192, 192, vp(b192.numpy()), None, None, None, 0, 0.0, None, None, 0, 0, 2)
    w32 = (A.double() @ Wq.double().t() + b192.double()).float()
    assert np.abs(out32 - w32.numpy()).max() <= 1e-5 * float(w32.abs().max())
    # ---- patch embedding: x = patches Wp^T + table[row % 197]; second output LN1(x) in bf16; mean / rstd per row
    Mp = 2 * 197
    patches = torch.randn(Mp, 768, generator=g).to(torch.bfloat16)
    Wp = (torch.randn(192, 768, generator=g) * 0.05).to(torch.bfloat16)
    table = torch.randn(197, 192, generator=g)
    gamma, beta = torch.randn(192, generator=g), torch.randn(192, generator=g)
    x = np.full((Mp, 192), np.nan, F)
    ln = np.full((Mp, 192), 0x7fc0, np.uint16)
    mean, rstd = np.full(Mp, np.nan, F), np.full(Mp, np.nan, F)
    gemm_nt_lib.gemm_nt(4, vp(_bf16_bits(patches)), vp(_bf16_bits(Wp)), vp(x), vp(ln), None, Mp, 192, 768, None, vp(gamma.numpy()), vp(beta.numpy()),
                        vp(table.numpy()), 197, 1e-6, vp(mean), vp(rstd), 1, 1, 2)
    xw = (patches.double() @ Wp.double().t() + table.double().repeat(2, 1)).float()
    assert np.abs(x - xw.numpy()).max() <= 1e-5 * float(xw.abs().max())
    lw = torch.nn.functional.layer_norm(xw, (192,), gamma, beta, 1e-6)
    assert float((_from_bits(ln) - lw).abs().max()) <= 2.0 ** -8 * float(lw.abs().max()) + 1e-4
    assert np.abs(mean - xw.mean(1).numpy()).max() <= 1e-5 and np.abs(rstd - (xw.var(1, unbiased=False) + 1e-6).rsqrt().numpy()).max() <= 1e-4


def test_forward_gemm_tcgen05_kernel_residual_by_tma_on_the_host(gemm_nt_lib):
    """EPI_RES_LN as the training blocks run it for fc2 / the attention projection: K = 768 (twelve K blocks through the 3-stage ring),
    the fp32 residual stream arrives as TMA-loaded panels in the staging slots, is rewritten in place with acc + bias + residual and
    leaves by TMA store; the next block's LayerNorm (bf16) and its row statistics come out of the same epilogue."""
    g = torch.Generator().manual_seed(2)
    M = 200
    H = torch.randn(M, 768, generator=g).to(torch.bfloat16)
    W2 = (torch.randn(192, 768, generator=g) * 0.05).to(torch.bfloat16)
    b2, gamma, beta = torch.randn(192, generator=g), torch.randn(192, generator=g), torch.randn(192, generator=g)
    res = torch.randn(M, 192, generator=g)
    x = np.full((M, 192), np.nan, F)
    ln = np.full((M, 192), 0x7fc0, np.uint16)
    mean, rstd = np.full(M, np.nan, F), np.full(M, np.nan, F)
    gemm_nt_lib.gemm_nt(4, vp(_bf16_bits(H)), vp(_bf16_bits(W2)), vp(x), vp(ln), vp(np.ascontiguousarray(res.numpy())), M, 192, 768, vp(b2.numpy()),
                        vp(gamma.numpy()), vp(beta.numpy()), None, 0, 1e-6, vp(mean), vp(rstd), 1, 1, 2)
    xw = (H.double() @ W2.double().t() + b2.double() + res.double()).float()
    assert np.abs(x - xw.numpy()).max() <= 1e-5 * float(xw.abs().max())
    lw = torch.nn.functional.layer_norm(xw, (192,), gamma, beta, 1e-6)
    assert float((_from_bits(ln) - lw).abs().max()) <= 2.0 ** -8 * float(lw.abs().max()) + 1e-4
    assert np.abs(mean - xw.mean(1).numpy()).max() <= 1e-5 and np.abs(rstd - (xw.var(1, unbiased=False) + 1e-6).rsqrt().numpy()).max() <= 1e-4


def test_forward_gemm_tcgen05_kernel_tiled_stream_on_the_host(gemm_nt_lib):
    """EPI_RES_LN on the INFERENCE path: the patch-embedding GEMM writes the fp32 residual stream in the tiled
    layout `[M/32][6 panels][8][32 rows][4]` straight from registers (no staging panels, no TMA) plus LayerNorm1 of block 0 in bf16;
    and the same mode with a tiled residual input (x' = acc + bias + x)."""
    g = torch.Generator().manual_seed(3)
    Mp = 2 * 197
    rows_pad = (Mp + 127) // 128 * 128
    patches = torch.randn(Mp, 768, generator=g).to(torch.bfloat16)
    Wp = (torch.randn(192, 768, generator=g) * 0.05).to(torch.bfloat16)
    table, gamma, beta = torch.randn(197, 192, generator=g), torch.randn(192, generator=g), torch.randn(192, generator=g)

    def untile(buf):           # xt_elem_offset (checked as a bijection in test_kernel_constants.py): [r/32][c/32][(c%32)/4][r%32][c%4]
        return buf.reshape(rows_pad // 32, 6, 8, 32, 4).transpose(0, 3, 1, 2, 4).reshape(rows_pad, 192)

    xt = np.full(rows_pad * 192, np.nan, F)
    ln = np.full((Mp, 192), 0x7fc0, np.uint16)
    gemm_nt_lib.gemm_nt_tiled(vp(_bf16_bits(patches)), vp(_bf16_bits(Wp)), vp(ln), Mp, 768, None, vp(gamma.numpy()), vp(beta.numpy()), vp(table.numpy()), 197,
                              None, vp(xt), 1e-6, 2)
    xw = (patches.double() @ Wp.double().t() + table.double().repeat(2, 1)).float()
    x = untile(xt)[:Mp]
    assert np.abs(x - xw.numpy()).max() <= 1e-5 * float(xw.abs().max())
    lw = torch.nn.functional.layer_norm(xw, (192,), gamma, beta, 1e-6)
    assert float((_from_bits(ln) - lw).abs().max()) <= 2.0 ** -8 * float(lw.abs().max()) + 1e-4
    # residual from the tiled stream itself, written back in place (x' = x + a W^T + b)
    A2 = torch.randn(Mp, 192, generator=g).to(torch.bfloat16)
    W2 = (torch.randn(192, 192, generator=g) * 0.1).to(torch.bfloat16)
    b2 = torch.randn(192, generator=g)
    xt2 = xt.copy()
    gemm_nt_lib.gemm_nt_tiled(vp(_bf16_bits(A2)), vp(_bf16_bits(W2)), vp(ln), Mp, 192, vp(b2.numpy()), vp(gamma.numpy()), vp(beta.numpy()), None, 0,
                              vp(xt2), vp(xt2), 1e-6, 2)
    xw2 = (torch.from_numpy(x.copy()).double() + A2.double() @ W2.double().t() + b2.double()).float()
    assert np.abs(untile(xt2)[:Mp] - xw2.numpy()).max() <= 1e-5 * float(xw2.abs().max())
