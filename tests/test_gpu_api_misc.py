"""API rows of SURVEY.md section 8a that had no test in round 1 (VERDICT rows A4, C6, weak #10) and the advisor's loss
findings: KANLayer.plot_activation, UncertaintyHead.sample, a non-default kan_layers model, fractional severity targets,
out-of-range class targets, non-integer focal gamma."""

import math

import numpy as np
import pytest
import torch

from conftest import assert_close
from oracle import kan as okan
from oracle import losses as olosses
from oracle import model as omodel

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from rovitkan_b200.models import RoViTKAN
    from rovitkan_b200.models.heads import UncertaintyHead
    from rovitkan_b200.models.kan import KANLayer
    from rovitkan_b200.training.losses import FocalLoss, JointLoss, KANRegressionLoss, OrdinalBCELoss, UncertaintyLoss

DEV = 'cuda'


def test_plot_activation_is_the_spline_of_one_edge():
    """kan.py:100-114: 100 points on [-1, 1], y = sum_k B_k(x) * W[i, j, k] (no tanh, no linear branch)."""
    torch.manual_seed(3)
    layer = KANLayer(6, 5).to(DEV)
    xs, ys = layer.plot_activation(input_idx=4, output_idx=2, num_points=100)
    assert isinstance(xs, np.ndarray) and xs.shape == (100,) and ys.shape == (100,)
    t = torch.linspace(-1, 1, 100)
    want = (okan.basis_literal(t[None], layer.knots.cpu(), 3)[0] * layer.spline_weights[4, 2].detach().cpu()).sum(1)
    assert_close(torch.from_numpy(xs), t, rtol=0, atol=1e-7, what='x grid')
    assert_close(torch.from_numpy(ys), want, rtol=1e-4, atol=1e-6, what='activation curve')
    assert float(np.abs(ys[xs >= 0.4 + 1e-6]).max()) == 0.0           # the truncated basis is dead beyond knots[7]
    assert torch.equal(layer.get_spline_weights(), layer.spline_weights.detach())


def test_uncertainty_head_sample_statistics():
    """heads.py:104-112: mu + exp(0.5*log_var) * eps, eps ~ N(0,1) of shape (B, S)."""
    torch.manual_seed(0)
    head = UncertaintyHead(embed_dim=192, hidden_dim=128, dropout=0.0).to(DEV).eval()
    x = torch.randn(16, 192, device=DEV)
    with torch.no_grad():
        mu, lv = head(x)
        s = head.sample(x, num_samples=20000)
    assert s.shape == (16, 20000)
    std = torch.exp(0.5 * lv)
    assert_close(s.mean(1, keepdim=True), mu, rtol=0, atol=4 * float(std.max()) / math.sqrt(20000), what='sample mean')
    assert_close(s.std(1, keepdim=True), std, rtol=3e-2, atol=0, what='sample std')


@pytest.mark.parametrize('kan_layers', [[192, 32, 1], [192, 48, 24, 8, 1]])
def test_non_default_kan_layers_model(kan_layers):
    """experiments/ablation.py varies kan_layers: such a model cannot use the one-kernel inference tail and must give the
    same numbers through the per-head kernels (eval and train), against the oracle at identical features."""
    sd = omodel.random_state_dict(2, kan_dims=kan_layers)
    m = RoViTKAN(pretrained=False, kan_layers=kan_layers, dropout=0.0)
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV).eval()
    assert m._fused_tail_params() is None
    x = torch.randn(6, 3, 224, 224, device=DEV)
    with torch.no_grad():
        o = m(x)
    sdd = {k: v.to(DEV) for k, v in sd.items()}
    ref = omodel.heads_forward(sdd, o['features'], 4)
    for k in ('cls_logits', 'ordinal_logits', 'mu', 'log_var', 'kan_severity'):
        assert_close(o[k], ref[k], rtol=1e-3, atol=1e-5, what=k)
    m.train()
    out = m(x)
    JointLoss()(out, torch.tensor([0, 1, 2, 3, 0, 1], device=DEV), torch.tensor([0, 1, 2, 3, 0, 1], device=DEV), 4)['total_loss'].backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
    # (train and eval trunks round differently, and the KAN is discontinuous: compare the tails at identical features)
    assert_close(m.kan_module(o['features']), o['kan_severity'], rtol=1e-5, atol=1e-6, what='train-mode KAN == eval-mode KAN')


def test_fractional_severity_targets_follow_the_reference_float_cast():
    """losses.py casts severities with .float(); 1.5 must not be truncated to 1 (NLL, MSE, and [y > k])."""
    torch.manual_seed(1)
    B = 37
    o = {'cls_logits': torch.randn(B, 4), 'ordinal_logits': torch.randn(B, 3), 'mu': torch.randn(B, 1) + 1.5,
         'log_var': torch.randn(B, 1), 'kan_severity': torch.rand(B, 1) * 3}
    yc = torch.randint(0, 4, (B,))
    ys = torch.rand(B) * 3                                        # continuous severities
    od = {k: v.to(DEV).requires_grad_(True) for k, v in o.items()}
    oc = {k: v.clone().requires_grad_(True) for k, v in o.items()}
    r = JointLoss()(od, yc.to(DEV), ys.to(DEV), 4)
    rr = olosses.joint(oc, yc, ys, 4)
    for k in rr:
        assert_close(r[k], rr[k], rtol=1e-4, atol=1e-6, what=k)
    r['total_loss'].backward()
    rr['total_loss'].backward()
    for k in o:
        assert_close(od[k].grad, oc[k].grad, rtol=1e-3, atol=1e-7, what='d/d' + k)
    assert_close(UncertaintyLoss()(od['mu'], od['log_var'], ys.to(DEV)), olosses.uncertainty_nll(o['mu'], o['log_var'], ys), rtol=1e-4, what='UncertaintyLoss')
    assert_close(KANRegressionLoss()(od['kan_severity'], ys.to(DEV)), olosses.kan_mse(o['kan_severity'], ys), rtol=1e-4, what='KANRegressionLoss')
    assert_close(OrdinalBCELoss()(od['ordinal_logits'], ys.to(DEV)), olosses.ordinal_bce(o['ordinal_logits'], ys), rtol=1e-4, what='OrdinalBCELoss')


def test_out_of_range_class_target_is_loud_and_never_indexes_out_of_bounds():
    logits = torch.randn(5, 4, device=DEV, requires_grad=True)
    y = torch.tensor([0, 1, 7, -100, 3], device=DEV)
    loss = FocalLoss()(logits, y)
    loss.backward()
    torch.cuda.synchronize()
    assert math.isnan(float(loss))
    assert torch.isfinite(logits.grad).all() and float(logits.grad[2].abs().sum()) == 0.0 and float(logits.grad[3].abs().sum()) == 0.0
    assert float(logits.grad[0].abs().sum()) > 0.0


def test_non_integer_focal_gamma_with_saturated_probability():
    """(1 - pt) can round below zero for a saturated logit; powf(negative, 1.5) would be NaN."""
    logits = torch.tensor([[40.0, -40.0, -40.0, -40.0], [0.3, 0.1, -0.2, 0.0]], device=DEV, requires_grad=True)
    y = torch.tensor([0, 2], device=DEV)
    loss = FocalLoss(gamma=1.5)(logits, y)
    loss.backward()
    lc = logits.detach().cpu().requires_grad_(True)
    ref = olosses.focal(lc, y.cpu(), gamma=1.5)
    ref.backward()
    assert_close(loss, ref, rtol=1e-4, atol=1e-7, what='focal gamma=1.5')
    assert torch.isfinite(logits.grad).all()
    assert_close(logits.grad[1], lc.grad[1], rtol=1e-3, atol=1e-7, what='focal gradient')


def test_hook_mode_attention_maps_and_probabilities():
    """SURVEY N4 / VERDICT B7: forward hooks on blocks[i].attn / .norm1 / .mlp / norm fire with the tensors of the fused
    trunk (the reference's explainability code registers exactly those: backbone.py:51-53, attention_maps.py:31-33,
    gradcam.py:40); checked against the oracle trunk's own intermediate tensors."""
    from oracle import vit as ovit
    sd = omodel.random_state_dict(7)
    m = RoViTKAN(pretrained=False, dropout=0.0)
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV).eval()
    x = torch.randn(3, 3, 224, 224, device=DEV)
    maps = m.get_attention_maps(x)                                   # rovit_kan.py:163-165 -> backbone.py:36-62
    assert len(maps) == 12 and all(t.shape == (3, 197, 192) for t in maps)
    # oracle intermediates
    trunk = ovit.DeiTTinyOracle()
    trunk.load_state_dict({k[len('backbone.model.'):]: v for k, v in sd.items() if k.startswith('backbone.model.')})
    trunk = trunk.to(DEV).eval()
    ref_attn, ref_norm1, ref_probs = [], [], []
    hooks = [b.attn.register_forward_hook(lambda mod, i, o: ref_attn.append(o)) for b in trunk.blocks]
    hooks += [b.norm1.register_forward_hook(lambda mod, i, o: ref_norm1.append(o)) for b in trunk.blocks]
    with torch.no_grad():
        f_ref = trunk(x)
    for h in hooks:
        h.remove()
    for i in (0, 5, 11):
        assert_close(maps[i], ref_attn[i], rtol=3e-2, atol=0, scale_tol=3e-2, what=f'blocks[{i}].attn output')
    got_norm1, got_final = [], []
    h1 = m.backbone.model.blocks[-1].norm1.register_forward_hook(lambda mod, i, o: got_norm1.append((i[0], o)))   # gradcam.py:40
    h2 = m.backbone.model.norm.register_forward_hook(lambda mod, i, o: got_final.append(o))
    with torch.no_grad():
        out = m(x)
    h1.remove(); h2.remove()
    assert len(got_norm1) == 1 and got_norm1[0][1].shape == (3, 197, 192)
    assert_close(got_norm1[0][1], ref_norm1[-1], rtol=3e-2, atol=0, scale_tol=3e-2, what='blocks[-1].norm1 output')
    assert torch.equal(got_final[0], out['features'])
    assert_close(out['features'], f_ref, rtol=2e-2, atol=0, scale_tol=2e-2, what='features in hook mode')
    with torch.no_grad():
        plain = m(x)                                                 # hooks removed: back on the fused inference path
    assert_close(plain['features'], out['features'], rtol=2e-2, atol=0, scale_tol=2e-2, what='hook mode vs fused path')
    probs = m.backbone.model.attention_probabilities(x)
    assert len(probs) == 12 and probs[0].shape == (3, 3, 197, 197)
    assert_close(probs[3].sum(-1), torch.ones(3, 3, 197, device=DEV), rtol=1e-5, atol=1e-5, what='rows of P sum to 1')
    # against softmax(q k^T / 8) recomputed from the oracle's qkv of that block
    blk = trunk.blocks[3]
    with torch.no_grad():
        qkv = blk.attn.qkv(ref_norm1[3]).reshape(3, 197, 3, 3, 64).permute(2, 0, 3, 1, 4)
        want = torch.softmax(qkv[0] @ qkv[1].transpose(-2, -1) * 0.125, dim=-1)
    assert_close(probs[3], want, rtol=5e-2, atol=2e-4, what='attention probabilities of block 3')


def test_kernel_protocol_error_surfaces_as_rovitkan_error():
    """every mbarrier wait of the library is bounded; a protocol error must end in a device-side trap that
    the host sees as RovitKanError (RVK_ERR_CUDA), not in a hung GPU.  Runs in its own process: the trap poisons the context."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, torch\n"
        f"sys.path.insert(0, {root!r})\n"
        "from rovitkan_b200 import _lib\n"
        "torch.cuda.init(); s = torch.cuda.current_stream().cuda_stream\n"
        "_lib.call('rvk_stream_check', s)\n"
        "_lib.call('rvk_debug_mbar_timeout', s)\n"
        "try:\n"
        "    _lib.call('rvk_stream_check', s)\n"
        "    print('NO ERROR')\n"
        "except _lib.RovitKanError as e:\n"
        "    print('SURFACED', e)\n")
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=120)
    assert 'SURFACED' in r.stdout and 'CUDA runtime error' in r.stdout, r.stdout + r.stderr
    assert 'mbarrier wait timed out' in r.stdout + r.stderr
    # the device is fine for the next process
    x = torch.ones(4, device=DEV)
    assert float(x.sum()) == 4.0


OUT_SHAPES = {'cls_logits': (4,), 'ordinal_logits': (3,), 'mu': (1,), 'log_var': (1,), 'kan_severity': (1,), 'features': (192,)}


@pytest.mark.parametrize('train', [False, True])
def test_empty_batch_gives_empty_outputs(train):
    """The reference's modules are shape-polymorphic in the batch: a 0-image batch (the tail of a drained loader) returns
    empty tensors of the right trailing shape, and its backward is a no-op that leaves zero gradients."""
    m = RoViTKAN(pretrained=False).to(DEV).train(train)
    x = torch.empty(0, 3, 224, 224, device=DEV)
    with torch.set_grad_enabled(train):
        o = m(x)
    for k, tail in OUT_SHAPES.items():
        assert tuple(o[k].shape) == (0,) + tail, (k, tuple(o[k].shape))
    if train:
        sum(o[k].sum() for k in OUT_SHAPES if k != 'features').backward()
        assert all(p.grad is None or float(p.grad.abs().max()) == 0.0 for p in m.parameters())
    else:
        p = m.predict(x)
        assert p['class'].shape == (0,) and tuple(p['class_probs'].shape) == (0, 4)
        assert tuple(p['ordinal_probs'].shape) == (0, 4) and p['ordinal_severity'].shape[0] == 0


def test_strided_and_channels_last_images_match_the_contiguous_call():
    """Loaders hand over channels_last / sliced batches; the kernels read packed NCHW, so the module must normalise the
    layout itself (the reference's Conv2d patch embedding accepts any strides)."""
    torch.manual_seed(1)
    m = RoViTKAN(pretrained=False).to(DEV).eval()
    big = torch.randn(7, 3, 224, 448, device=DEV)
    x = big[:, :, :, ::2]                                    # strided view
    assert not x.is_contiguous()
    with torch.no_grad():
        ref = m(x.contiguous())
        o1 = m(x)
        o2 = m(x.contiguous().to(memory_format=torch.channels_last))
    for k in OUT_SHAPES:
        assert torch.equal(o1[k], ref[k]), k
        assert torch.equal(o2[k], ref[k]), k
