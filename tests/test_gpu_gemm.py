"""GPU parity of the tcgen05/TMEM GEMM kernels, called through the C ABI, against fp32 matmuls of the
same bf16-rounded operands.  Tolerances: outputs stored as bf16 carry one bf16 rounding (2^-9
relative); fp32 outputs only differ by accumulation order."""

import pytest
import torch

from conftest import assert_close

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from rovitkan_b200 import _lib

DEV = 'cuda'


def _s():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return 0 if t is None else t.data_ptr()


def _rand_bf16(*shape, seed=0, scale=1.0):
    g = torch.Generator(device='cpu').manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV).to(torch.bfloat16)


def gemm_nt(mode, A, B, out, out2=None, aux=None, bias=None, gamma=None, beta=None, table=None, eps=1e-6,
            mean=None, rstd=None):
    M, K = A.shape
    N = B.shape[0]
    _lib.call('rvk_gemm_nt', mode, _p(A), A.stride(0), _p(B), B.stride(0), _p(out), out.stride(0), _p(out2),
              0 if out2 is None else out2.stride(0), _p(aux), 0 if aux is None else aux.stride(0), M, N, K, _p(bias),
              _p(gamma), _p(beta), _p(table), 0 if table is None else table.shape[0], eps, _p(mean), _p(rstd), _s())
    torch.cuda.synchronize()


def gelu(x):
    return 0.5 * x * (1 + torch.erf(x / 2 ** 0.5))


def gelu_grad(x):
    return 0.5 * (1 + torch.erf(x / 2 ** 0.5)) + x * torch.exp(-0.5 * x * x) / (2 * torch.pi) ** 0.5


SHAPES = [(128, 192, 192), (300, 576, 192), (1000, 768, 192), (777, 192, 768), (19 * 128 + 5, 576, 192),
          (40000, 192, 192)]


@pytest.mark.parametrize('M,N,K', SHAPES)
def test_gemm_bias_bf16(M, N, K):
    A, B = _rand_bf16(M, K, seed=1), _rand_bf16(N, K, seed=2, scale=0.1)
    bias = torch.randn(N, device=DEV)
    out = torch.full((M, N), float('nan'), device=DEV, dtype=torch.bfloat16)
    gemm_nt(0, A, B, out, bias=bias)
    ref = A.float() @ B.float().t() + bias
    assert_close(out.float(), ref, rtol=8e-3, atol=2e-3, what=f'gemm bf16 {M}x{N}x{K}')


def test_gemm_no_bias_and_strided_views():
    M, N, K = 500, 192, 576
    A_full = _rand_bf16(M, 768, seed=3)
    A = A_full[:, 64:64 + K]                      # leading dimension 768, 128-byte aligned column offset
    B = _rand_bf16(N, K, seed=4, scale=0.1)
    out = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
    gemm_nt(0, A, B, out)
    assert_close(out.float(), A.float() @ B.float().t(), rtol=8e-3, atol=2e-3, what='strided A')


@pytest.mark.parametrize('M', [128, 1000, 5000])
def test_gemm_f32_out(M):
    N, K = 576, 192
    A, B = _rand_bf16(M, K, seed=5), _rand_bf16(N, K, seed=6, scale=0.1)
    bias = torch.randn(N, device=DEV)
    out = torch.full((M, N), float('nan'), device=DEV)
    gemm_nt(3, A, B, out, bias=bias)
    ref = (A.double() @ B.double().t() + bias.double()).float()
    assert_close(out, ref, rtol=1e-4, atol=1e-4, what='gemm fp32 out')


@pytest.mark.parametrize('M,with_z', [(256, False), (1000, True), (9000, True)])
def test_gemm_gelu(M, with_z):
    N, K = 768, 192
    A, B = _rand_bf16(M, K, seed=7), _rand_bf16(N, K, seed=8, scale=0.1)
    bias = torch.randn(N, device=DEV) * 0.5
    h = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
    z = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16) if with_z else None
    gemm_nt(1, A, B, h, out2=z, bias=bias)
    zr = A.float() @ B.float().t() + bias
    assert_close(h.float(), gelu(zr), rtol=8e-3, atol=3e-3, what='gelu(h)')
    if with_z:
        assert_close(z.float(), zr, rtol=8e-3, atol=3e-3, what='z')


@pytest.mark.parametrize('M', [200, 4096])
def test_gemm_dgelu(M):
    N, K = 768, 192
    A, B = _rand_bf16(M, K, seed=9), _rand_bf16(N, K, seed=10, scale=0.1)
    z = _rand_bf16(M, N, seed=11, scale=1.5)
    out = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
    gemm_nt(2, A, B, out, aux=z)
    ref = (A.float() @ B.float().t()) * gelu_grad(z.float())
    assert_close(out.float(), ref, rtol=1e-2, atol=3e-3, what='dgelu')


@pytest.mark.parametrize('M,K,res,ln', [(128, 192, 'tma', True), (1000, 768, 'tma', True), (197 * 3, 768, 'table', True),
                                        (2500, 192, 'tma', False), (700, 192, 'none', True), (197 * 64, 192, 'tma', True)])
def test_gemm_residual_layernorm(M, K, res, ln):
    N = 192
    A, B = _rand_bf16(M, K, seed=12), _rand_bf16(N, K, seed=13, scale=0.1)
    bias = torch.randn(N, device=DEV) * 0.3
    gamma = 1 + 0.2 * torch.randn(N, device=DEV)
    beta = 0.1 * torch.randn(N, device=DEV)
    x_old = torch.randn(M, N, device=DEV) * 2 if res == 'tma' else None
    table = torch.randn(197, N, device=DEV) if res == 'table' else None
    x_new = torch.full((M, N), float('nan'), device=DEV)
    y = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16) if ln else None
    mean = torch.zeros(M, device=DEV) if ln else None
    rstd = torch.zeros(M, device=DEV) if ln else None
    gemm_nt(4, A, B, x_new, out2=y, aux=x_old, bias=bias, gamma=gamma if ln else None, beta=beta if ln else None,
            table=table, mean=mean, rstd=rstd)
    ref = (A.double() @ B.double().t()).float() + bias
    if res == 'tma':
        ref = ref + x_old
    elif res == 'table':
        ref = ref + table[torch.arange(M, device=DEV) % 197]
    assert_close(x_new, ref, rtol=1e-4, atol=2e-4, what='x_new')
    if ln:
        mu = ref.mean(1)
        var = ref.var(1, unbiased=False)
        assert_close(mean, mu, rtol=1e-4, atol=1e-4, what='mean')
        assert_close(rstd, torch.rsqrt(var + 1e-6), rtol=1e-4, atol=1e-5, what='rstd')
        yr = torch.nn.functional.layer_norm(ref, (N,), gamma, beta, 1e-6)
        assert_close(y.float(), yr, rtol=8e-3, atol=4e-3, what='layernorm out')


def test_gemm_residual_in_place():
    M, K, N = 1500, 192, 192
    A, B = _rand_bf16(M, K, seed=14), _rand_bf16(N, K, seed=15, scale=0.1)
    x = torch.randn(M, N, device=DEV)
    ref = x + (A.double() @ B.double().t()).float()
    gemm_nt(4, A, B, x, aux=x)
    assert_close(x, ref, rtol=1e-4, atol=2e-4, what='in-place residual')


@pytest.mark.parametrize('M,P,Q', [(64, 192, 192), (1000, 192, 768), (5000, 768, 192), (4097, 576, 192),
                                   (197 * 256, 192, 768)])
def test_gemm_tn_weight_gradient(M, P, Q):
    A, B = _rand_bf16(M, P, seed=16), _rand_bf16(M, Q, seed=17)
    C = torch.randn(P, Q, device=DEV)
    ref = C.double() + 0.5 * (A.double().t() @ B.double())
    fused_bias = Q <= 192                       # bias gradient (column sums of A) through the ones column of the same GEMM
    cs = torch.full((P,), 0.25, device=DEV)
    _lib.call('rvk_gemm_tn', _p(A), A.stride(0), _p(B), B.stride(0), _p(C), C.stride(0), M, P, Q, 0.5,
              _p(cs) if fused_bias else 0, _s())
    torch.cuda.synchronize()
    assert_close(C, ref.float(), rtol=1e-4, atol=1e-4, scale_tol=2e-5, what=f'wgrad {M}x{P}x{Q}')
    if fused_bias:
        assert_close(cs, (0.25 + 0.5 * A.double().sum(0)).float(), rtol=1e-4, atol=1e-3, scale_tol=2e-5, what='fused column sums of A')


def test_gemm_rejects_bad_shapes():
    A, B = _rand_bf16(128, 192), _rand_bf16(100, 192)
    out = torch.zeros(128, 100, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(_lib.RovitKanError):
        gemm_nt(0, A, B, out)
