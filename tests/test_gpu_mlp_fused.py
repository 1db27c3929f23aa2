"""GPU parity of the fused MLP block kernel (fc1 + GELU + fc2 + residual + LayerNorm in one tcgen05 kernel, the
LayerNorm applied on load, hidden activation never leaving the SM) against an fp32 restatement of timm's Block MLP half on the same
bf16-rounded operands (oracle/vit.py::_Block).  The CTA-pair (cta_group::2) kernel, its single-CTA variant (G = 1) and the
two-row-tiles-in-flight kernel the inference path runs by default (G = 4, csrc/mlp_fused2.cuh; needs the folded projection).
Tolerances: GELU is evaluated in packed fp16 arithmetic and the hidden activation is rounded to fp16 before fc2
(together about one bf16 rounding, 2^-9 relative, per hidden element -- the unfused path rounds it to bf16);
x_out is fp32, ln_out carries one more bf16 rounding."""

import pytest
import torch

from conftest import assert_close

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from rovitkan_b200 import _lib

DEV = 'cuda'


def to_tiled(x):
    """[M,192] fp32 -> tiled token stream (include/rovitkan.h, rvk_mlp_fused), rows padded to 128."""
    M = x.shape[0]
    Mp = (M + 127) // 128 * 128
    xp = torch.zeros(Mp, 192, device=x.device, dtype=x.dtype)
    xp[:M] = x
    return xp.view(Mp // 32, 32, 6, 8, 4).permute(0, 2, 3, 1, 4).contiguous()


def from_tiled(t, M):
    return t.permute(0, 3, 1, 2, 4).reshape(-1, 192)[:M]


def gelu(x):
    return 0.5 * x * (1 + torch.erf(x / 2 ** 0.5))


def run_case(M, G, has_ln, seed=0, inplace=False, proj=False):
    g = torch.Generator().manual_seed(seed)
    w1 = (torch.randn(768, 192, generator=g) * 0.08).to(DEV).to(torch.bfloat16)
    w2 = (torch.randn(192, 768, generator=g) * 0.05).to(DEV).to(torch.float16)
    b1 = (torch.randn(768, generator=g) * 0.5).to(DEV)
    b2 = torch.randn(192, generator=g).to(DEV)
    gamma2 = (1 + 0.2 * torch.randn(192, generator=g)).to(DEV)
    beta2 = (0.3 * torch.randn(192, generator=g)).to(DEV)
    gamma = (1 + 0.2 * torch.randn(192, generator=g)).to(DEV)
    beta = (0.3 * torch.randn(192, generator=g)).to(DEV)
    x = (2.0 * torch.randn(M, 192, generator=g) + 0.5).to(DEV)
    xt = to_tiled(x)
    xo = xt if inplace else torch.full_like(xt, float('nan'))
    ln_out = torch.full((M, 192), float('nan'), device=DEV, dtype=torch.bfloat16) if has_ln else None
    tail = (gamma2.data_ptr(), beta2.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
            gamma.data_ptr() if has_ln else 0, beta.data_ptr() if has_ln else 0, 1e-6, ln_out.data_ptr() if has_ln else 0, M, G,
            torch.cuda.current_stream().cuda_stream)
    if proj:
        # attention output projection folded in: x <- x + ctx . Wp^T + bp before the MLP half (fp32 accumulation of bf16 operands)
        ctx = torch.randn(M, 192, generator=g).to(DEV).to(torch.bfloat16)
        wp = (torch.randn(192, 192, generator=g) * 0.08).to(DEV).to(torch.bfloat16)
        bp = torch.randn(192, generator=g).to(DEV)
        _lib.call('rvk_attn_proj_mlp_fused', xt.data_ptr(), xo.data_ptr(), ctx.data_ptr(), wp.data_ptr(), bp.data_ptr(), *tail)
        x = x + ctx.float() @ wp.float().t() + bp
    else:
        _lib.call('rvk_mlp_fused', xt.data_ptr(), xo.data_ptr(), *tail)
    torch.cuda.synchronize()
    a = torch.nn.functional.layer_norm(x, (192,), gamma2, beta2, 1e-6).to(torch.bfloat16).float()   # A operand is bf16
    h = gelu(a @ w1.float().t() + b1)      # kept in fp16 on chip (GELU itself evaluated in half2)
    ref = x + h @ w2.float().t() + b2
    got = from_tiled(xo, M)
    assert_close(got, ref, rtol=2e-3, atol=2e-3, scale_tol=1e-3, what=f'mlp_fused x_out M={M} G={G} proj={proj}')
    if has_ln:
        ref_ln = torch.nn.functional.layer_norm(ref, (192,), gamma, beta, 1e-6)
        assert_close(ln_out.float(), ref_ln, rtol=1e-2, atol=1e-2, what=f'mlp_fused ln_out M={M} G={G}')
        assert bool(torch.isfinite(ln_out.float()).all())


@pytest.mark.parametrize('M', [128, 300, 1000, 197 * 64])
def test_mlp_fused_single_cta(M):
    run_case(M, 1, True, seed=M)


@pytest.mark.parametrize('M', [128, 256, 300, 1000, 197 * 64, 197 * 300 + 5])
def test_mlp_fused_cta_pair(M):
    run_case(M, 2, True, seed=M + 1)


@pytest.mark.parametrize('G', [1, 2])
def test_mlp_fused_no_layernorm_in_place(G):
    run_case(1000, G, False, seed=7, inplace=True)
    run_case(197 * 40, G, True, seed=8, inplace=True)


@pytest.mark.parametrize('G', [1, 2, 4])
@pytest.mark.parametrize('M', [128, 300, 1000, 197 * 64, 197 * 300 + 5])
def test_attn_proj_mlp_fused(M, G):
    """rvk_attn_proj_mlp_fused: x + proj(ctx) on the tensor cores (G = 1, 2: in the idle TMEM columns; G = 4: in the tile's own
    accumulator slot, where the projected row then stays parked under fc2), then the MLP half on the new rows."""
    run_case(M, G, True, seed=M + 11, proj=True)
    run_case(M, G, M % 2 == 0, seed=M + 12, inplace=True, proj=True)


@pytest.mark.parametrize('G', [2, 4])
def test_attn_proj_mlp_fused_equals_two_launches_at_benchmark_size(G):
    """BASELINE configs[1] size (1024 images = 201 728 token rows): folding the attention output projection into the MLP kernel
    must give what the separate projection GEMM followed by rvk_mlp_fused gives (same bf16 operands, fp32 accumulation; the
    only difference is the order of two fp32 additions per element)."""
    M = 1024 * 197
    g = torch.Generator().manual_seed(99)
    mk = lambda *sh, sc=1.0: (torch.randn(*sh, generator=g) * sc).to(DEV)
    w1, w2 = mk(768, 192, sc=0.08).to(torch.bfloat16), mk(192, 768, sc=0.05).to(torch.float16)
    b1, b2, bp = mk(768, sc=0.5), mk(192), mk(192)
    gamma2, beta2, gamma, beta = 1 + mk(192, sc=0.2), mk(192, sc=0.3), 1 + mk(192, sc=0.2), mk(192, sc=0.3)
    wp = mk(192, 192, sc=0.08).to(torch.bfloat16)
    x = to_tiled(mk(M, 192, sc=2.0) + 0.5)
    ctx = mk(M, 192).to(torch.bfloat16)
    s = torch.cuda.current_stream().cuda_stream
    tail = (gamma2.data_ptr(), beta2.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), gamma.data_ptr(),
            beta.data_ptr(), 1e-6)
    xa, lna = x.clone(), torch.empty(M, 192, device=DEV, dtype=torch.bfloat16)
    _lib.call('rvk_attn_proj_mlp_fused', xa.data_ptr(), xa.data_ptr(), ctx.data_ptr(), wp.data_ptr(), bp.data_ptr(), *tail,
              lna.data_ptr(), M, G, s)
    # two launches: x += ctx . Wp^T + bp through the fp32 GEMM epilogue (row-major round trip), then the MLP kernel
    xr = from_tiled(x, M).contiguous()
    t = torch.empty(M, 192, device=DEV)
    _lib.call('rvk_gemm_nt', 3, ctx.data_ptr(), 192, wp.data_ptr(), 192, t.data_ptr(), 192, 0, 0, 0, 0, M, 192, 192, bp.data_ptr(),
              0, 0, 0, 0, 1e-6, 0, 0, s)
    xb = to_tiled(xr + t)
    lnb = torch.empty(M, 192, device=DEV, dtype=torch.bfloat16)
    _lib.call('rvk_mlp_fused', xb.data_ptr(), xb.data_ptr(), *tail, lnb.data_ptr(), M, 2, s)
    torch.cuda.synchronize()
    assert_close(from_tiled(xa, M), from_tiled(xb, M), rtol=2e-3, atol=2e-3, scale_tol=1e-3, what='fused vs two launches (x)')
    assert float((lna.float() - lnb.float()).abs().mean()) < 2e-3
