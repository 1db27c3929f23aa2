"""GPU parity of the non-GEMM encoder kernels (attention fwd/bwd, LayerNorm fwd/bwd, im2col), called
through the C ABI, against fp32 PyTorch restatements of the same op on the same bf16-rounded inputs."""

import pytest
import torch

from conftest import assert_close

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from rovitkan_b200 import _lib

DEV = 'cuda'
TOK = 197


def _s():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return 0 if t is None else t.data_ptr()


def _attention_ref(qkv):
    b = qkv.shape[0] // TOK
    q, k, v = qkv.float().reshape(b, TOK, 3, 3, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * 0.125
    p = torch.softmax(s, dim=-1)
    ctx = (p @ v).transpose(1, 2).reshape(b * TOK, 192)
    lse2 = torch.logsumexp(s, dim=-1) * 1.4426950408889634       # kernel stores log2-domain lse
    return ctx, lse2


@pytest.mark.parametrize('batch', [1, 3, 64])
def test_attention_forward(batch):
    g = torch.Generator().manual_seed(0)
    qkv = (torch.randn(batch * TOK, 576, generator=g) * 1.5).to(DEV).to(torch.bfloat16)
    ctx = torch.full((batch * TOK, 192), float('nan'), device=DEV, dtype=torch.bfloat16)
    lse = torch.zeros(batch, 3, TOK, device=DEV)
    _lib.call('rvk_attention_forward', _p(qkv), _p(ctx), _p(lse), batch, _s())
    torch.cuda.synchronize()
    ref, lse_ref = _attention_ref(qkv)
    # P is rounded to bf16 before P*V and the output is stored as bf16
    assert_close(ctx.float(), ref, rtol=1e-2, atol=1e-2, what='attention ctx')
    assert_close(lse, lse_ref, rtol=1e-4, atol=1e-3, what='lse')


@pytest.mark.parametrize('batch', [1, 5, 60])      # 60 images = 180 (image, head) items: CTAs that walk more than one item
def test_attention_backward(batch):
    g = torch.Generator().manual_seed(1)
    qkv = (torch.randn(batch * TOK, 576, generator=g)).to(DEV).to(torch.bfloat16)
    dctx = (torch.randn(batch * TOK, 192, generator=g)).to(DEV).to(torch.bfloat16)
    ctx = torch.zeros(batch * TOK, 192, device=DEV, dtype=torch.bfloat16)
    lse = torch.zeros(batch, 3, TOK, device=DEV)
    _lib.call('rvk_attention_forward', _p(qkv), _p(ctx), _p(lse), batch, _s())
    dqkv = torch.full((batch * TOK, 576), float('nan'), device=DEV, dtype=torch.bfloat16)
    _lib.call('rvk_attention_backward', _p(qkv), _p(ctx), _p(dctx), _p(lse), _p(dqkv), batch, _s())
    torch.cuda.synchronize()
    x = qkv.float().requires_grad_(True)
    q, k, v = x.reshape(batch, TOK, 3, 3, 64).permute(2, 0, 3, 1, 4)
    p = torch.softmax((q @ k.transpose(-1, -2)) * 0.125, dim=-1)
    ref_ctx = (p @ v).transpose(1, 2).reshape(batch * TOK, 192)
    ref_ctx.backward(dctx.float())
    assert_close(dqkv.float(), x.grad, rtol=2e-2, atol=2e-2, scale_tol=5e-3, what='dqkv')


def test_attention_backward_properties_at_training_batch():
    """Size-independent properties at BASELINE configs[3]'s 256 images per GPU: an image's gradient does not depend on the batch it
    travels in (bitwise: items are independent), and the backward is linear in the upstream gradient."""
    batch = 256
    g = torch.Generator().manual_seed(2)
    qkv = (torch.randn(batch * TOK, 576, generator=g)).to(DEV).to(torch.bfloat16)
    dctx = (torch.randn(batch * TOK, 192, generator=g)).to(DEV).to(torch.bfloat16)

    def run(qkv_, dctx_, n):
        ctx = torch.zeros(n * TOK, 192, device=DEV, dtype=torch.bfloat16)
        lse = torch.zeros(n, 3, TOK, device=DEV)
        _lib.call('rvk_attention_forward', _p(qkv_), _p(ctx), _p(lse), n, _s())
        out = torch.full((n * TOK, 576), float('nan'), device=DEV, dtype=torch.bfloat16)
        _lib.call('rvk_attention_backward', _p(qkv_), _p(ctx), _p(dctx_), _p(lse), _p(out), n, _s())
        torch.cuda.synchronize()
        return out
    full = run(qkv, dctx, batch)
    assert bool(torch.isfinite(full.float()).all())
    for img in (0, 147, 255):
        one = run(qkv[img * TOK:(img + 1) * TOK].contiguous(), dctx[img * TOK:(img + 1) * TOK].contiguous(), 1)
        assert torch.equal(one, full[img * TOK:(img + 1) * TOK]), f'image {img}: gradient depends on the batch'
    twice = run(qkv, (dctx.float() * 2).to(torch.bfloat16), batch)          # exact in bf16: a power of two
    assert_close(twice.float(), 2 * full.float(), rtol=1e-2, atol=1e-3, what='linearity in dctx')


@pytest.mark.parametrize('rows,bf16', [(1, True), (37, False), (5000, True)])
def test_layernorm_forward(rows, bf16):
    x = torch.randn(rows, 192, device=DEV) * 3 + 0.5
    gamma, beta = torch.randn(192, device=DEV), torch.randn(192, device=DEV)
    y = torch.zeros(rows, 192, device=DEV, dtype=torch.bfloat16 if bf16 else torch.float32)
    mean, rstd = torch.zeros(rows, device=DEV), torch.zeros(rows, device=DEV)
    _lib.call('rvk_layernorm_forward', _p(x), 192, _p(gamma), _p(beta), 1e-6, _p(y), int(bf16), 192, _p(mean), _p(rstd),
              rows, _s())
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm(x, (192,), gamma, beta, 1e-6)
    assert_close(y.float(), ref, rtol=8e-3 if bf16 else 1e-5, atol=8e-3 if bf16 else 1e-5, what='ln fwd')
    assert_close(mean, x.mean(1), rtol=1e-5, atol=1e-5, what='mean')
    assert_close(rstd, torch.rsqrt(x.var(1, unbiased=False) + 1e-6), rtol=1e-5, atol=1e-6, what='rstd')


def test_layernorm_forward_strided_cls_rows():
    b = 9
    x = torch.randn(b * TOK, 192, device=DEV)
    gamma, beta = torch.randn(192, device=DEV), torch.randn(192, device=DEV)
    y = torch.zeros(b, 192, device=DEV)
    _lib.call('rvk_layernorm_forward', _p(x), TOK * 192, _p(gamma), _p(beta), 1e-6, _p(y), 0, 192, 0, 0, b, _s())
    torch.cuda.synchronize()
    assert_close(y, torch.nn.functional.layer_norm(x[::TOK], (192,), gamma, beta, 1e-6), rtol=1e-5, atol=1e-5,
                 what='cls-row ln')


@pytest.mark.parametrize('rows,g_bf16', [(3, False), (17, True), (4000, True), (50432, False)])
def test_layernorm_backward(rows, g_bf16):
    x = (torch.randn(rows, 192, device=DEV) * 2).requires_grad_(True)
    gamma = torch.randn(192, device=DEV, requires_grad=True)
    beta = torch.randn(192, device=DEV, requires_grad=True)
    g = torch.randn(rows, 192, device=DEV)
    if g_bf16:
        g = g.to(torch.bfloat16)
    dx_in = torch.randn(rows, 192, device=DEV)
    y = torch.nn.functional.layer_norm(x, (192,), gamma, beta, 1e-6)
    y.backward(g.float())
    mean = x.detach().mean(1)
    rstd = torch.rsqrt(x.detach().var(1, unbiased=False) + 1e-6)
    dx = dx_in.clone()
    dxb = torch.zeros(rows, 192, device=DEV, dtype=torch.bfloat16)
    dgamma, dbeta = torch.zeros(192, device=DEV), torch.zeros(192, device=DEV)
    dcol = torch.full((192,), 0.5, device=DEV)          # accumulates: column sums of the dx written (fused bias gradient)
    _lib.call('rvk_layernorm_backward', _p(g), int(g_bf16), 192, _p(x.detach()), 192, _p(mean), _p(rstd),
              _p(gamma.detach()), _p(dx), _p(dx), 192, _p(dxb), _p(dgamma), _p(dbeta), _p(dcol), rows, _s())
    torch.cuda.synchronize()
    assert_close(dcol, 0.5 + dx.sum(0), rtol=1e-3, atol=1e-3, scale_tol=1e-4, what='fused column sum of dx')
    assert_close(dx, dx_in + x.grad, rtol=1e-4, atol=1e-4, what='dx')
    assert_close(dxb.float(), dx, rtol=8e-3, atol=1e-3, what='dx bf16 copy')
    assert_close(dgamma, gamma.grad, rtol=1e-3, atol=1e-3, scale_tol=1e-4, what='dgamma')
    assert_close(dbeta, beta.grad, rtol=1e-3, atol=1e-3, scale_tol=1e-4, what='dbeta')


@pytest.mark.parametrize('batch', [1, 7])
def test_im2col(batch):
    img = torch.randn(batch, 3, 224, 224, device=DEV)
    out = torch.full((batch * TOK, 768), float('nan'), device=DEV, dtype=torch.bfloat16)
    _lib.call('rvk_im2col', _p(img), _p(out), batch, _s())
    torch.cuda.synchronize()
    patches = img.reshape(batch, 3, 14, 16, 14, 16).permute(0, 2, 4, 1, 3, 5).reshape(batch, 196, 768)
    ref = torch.cat([torch.zeros(batch, 1, 768, device=DEV), patches], dim=1).reshape(batch * TOK, 768)
    assert torch.equal(out, ref.to(torch.bfloat16))
