"""Pin the CPU oracle against outputs of the reference itself (tests/golden/*.npz,
made by tests/golden/make_golden.py from /root/reference) and against the
RNG-free known-answer tables of SURVEY.md section 4.  CPU only."""

import numpy as np
import pytest
import torch

from conftest import T, assert_close, load_golden
from oracle import heads as oheads
from oracle import kan as okan
from oracle import losses as olosses
from oracle import model as omodel
from oracle import vit as ovit


# ----------------------------------------------------------------------------- KAN basis
KAT_BASIS = {   # rows of BSplineBasis.compute_basis at t
    -1.0: [0, 0, 0, 0, 0, 0, 0],
    -0.9: [.0208333, 0, 0, 0, 0, 0, 0],
    -0.8: [.1666667, 0, 0, 0, 0, 0, 0],
    -0.7: [.4791667, .0208333, 0, 0, 0, 0, 0],
    -0.5: [.4791667, .4791667, .0208333, 0, 0, 0, 0],
    -0.3: [.0208333, .4791666, .4791667, .0208333, 0, 0, 0],
    0.0: [0, 0, .1666667, .6666667, .1666666, 0, 0],
    0.3: [0, 0, 0, .0208333, .4791667, .4791667, .0208333],
    0.4: [0] * 7, 0.5: [0] * 7, 0.7: [0] * 7, 1.0: [0] * 7,
}


@pytest.mark.parametrize('fn', [okan.basis_literal, okan.basis_closed_form])
def test_basis_known_answers(fn):
    knots = okan.make_knots()
    for t, row in KAT_BASIS.items():
        got = fn(torch.tensor([[t]], dtype=torch.float32), knots)[0, 0]
        assert_close(got, torch.tensor(row, dtype=torch.float32), rtol=0, atol=2e-6, what=f'basis({t})')


@pytest.mark.parametrize('fn', [okan.basis_literal, okan.basis_closed_form])
def test_basis_matches_reference_vectors(fn):
    g = load_golden('kan_basis.npz')
    got = fn(T(g['t']), T(g['knots']))
    tol = 0 if fn is okan.basis_literal else 1e-6
    assert_close(got, T(g['basis']), rtol=0, atol=tol, what=fn.__name__)


def test_basis_dead_zone_and_partition_of_unity():
    knots = okan.make_knots()
    t = torch.linspace(-1, 1, 2001)[None]
    b = okan.basis_literal(t, knots)[0]
    dead = t[0] >= knots[7]
    assert float(b[dead].abs().max()) == 0.0
    inner = (t[0] >= knots[3]) & (t[0] < knots[7])
    assert_close(b[inner].sum(-1), torch.ones(int(inner.sum())), rtol=0, atol=1e-6, what='partition of unity')


def test_closed_form_derivative_matches_autograd():
    knots = okan.make_knots()
    t = (torch.rand(1, 513, dtype=torch.float64) * 1.6 - 1.0)
    t = t[(t - t.round(decimals=1)).abs() > 1e-3][None]          # stay off the knots
    t32 = t.float().requires_grad_(True)
    for k in range(7):
        val = okan.basis_literal(t32, knots)[0, :, k].sum()
        (grad,) = torch.autograd.grad(val, t32)
        _, d = okan.basis_closed_form(t32.detach(), knots, with_derivative=True)
        assert_close(d[0, :, k], grad[0], rtol=1e-4, atol=1e-4, what=f"N'_{k}")


# ----------------------------------------------------------------------------- KAN layers
@pytest.mark.parametrize('tag', ['l0', 'l1', 'l2', 'odd'])
@pytest.mark.parametrize('loop', [False, True])
def test_kan_layer_matches_reference(tag, loop):
    g = load_golden('kan_layers.npz')
    if loop and tag == 'l0':
        pytest.skip('12k-iteration loop covered by the smaller layers')
    x = T(g[f'{tag}_x']).requires_grad_(True)
    sw, lw, lb = (T(g[f'{tag}_{k}']).requires_grad_(True) for k in ('sw', 'lw', 'lb'))
    fn = okan.layer_forward_loop if loop else okan.layer_forward
    y = fn(x, sw, lw, lb, T(g[f'{tag}_knots']))
    assert_close(y, T(g[f'{tag}_y']), rtol=1e-5, atol=2e-6, what='y')
    y.backward(T(g[f'{tag}_gy']))
    assert_close(x.grad, T(g[f'{tag}_dx']), rtol=1e-4, atol=1e-5, what='dx')
    assert_close(sw.grad, T(g[f'{tag}_dsw']), rtol=1e-4, atol=1e-5, what='dW')
    assert_close(lw.grad, T(g[f'{tag}_dlw']), rtol=1e-4, atol=1e-5, what='dWl')
    assert_close(lb.grad, T(g[f'{tag}_dlb']), rtol=1e-4, atol=1e-5, what='db')


def test_kan_module_matches_reference():
    g = load_golden('kan_layers.npz')
    layers = [tuple(T(g[f'mod_{k}{i}']).requires_grad_(True) for k in ('sw', 'lw', 'lb')) for i in range(3)]
    x = T(g['mod_x']).requires_grad_(True)
    traj = okan.severity_forward(x, layers, okan.make_knots(), return_trajectory=True)
    for i, a in enumerate(traj):
        assert_close(a, T(g[f'mod_traj{i}']), rtol=1e-5, atol=2e-6, what=f'trajectory {i}')
    traj[-1].backward(T(g['mod_gy']))
    # atol covers fp32 re-association between the reference's 12k-step loop and the einsum
    assert_close(x.grad, T(g['mod_dx']), rtol=1e-4, atol=2e-5, what='dx')
    for i, (sw, lw, lb) in enumerate(layers):
        assert_close(sw.grad, T(g[f'mod_dsw{i}']), rtol=1e-4, atol=2e-5, what=f'dW{i}')
        assert_close(lw.grad, T(g[f'mod_dlw{i}']), rtol=1e-4, atol=2e-5, what=f'dWl{i}')
        assert_close(lb.grad, T(g[f'mod_dlb{i}']), rtol=1e-4, atol=2e-5, what=f'db{i}')
    assert sum(t.numel() for l in layers for t in l) == 106705          # README.md:316


# ----------------------------------------------------------------------------- heads
def test_heads_match_reference():
    g = load_golden('heads.npz')
    x = T(g['x']).requires_grad_(True)
    P = {k: T(g[k]).requires_grad_(True) for k in g.files if '.' in k and not k.endswith('.grad')}
    cls = oheads.classification_forward(x, P['cls.fc1.weight'], P['cls.fc1.bias'], P['cls.fc2.weight'], P['cls.fc2.bias'])
    ordl = oheads.ordinal_forward(x, P['ord.fc1.weight'], P['ord.fc1.bias'], P['ord.fc2.weight'], P['ord.fc2.bias'])
    mu, lv = oheads.uncertainty_forward(x, P['unc.fc1.weight'], P['unc.fc1.bias'], P['unc.fc_mu.weight'],
                                        P['unc.fc_mu.bias'], P['unc.fc_logvar.weight'], P['unc.fc_logvar.bias'])
    for got, key in ((cls, 'cls'), (ordl, 'ord'), (mu, 'mu'), (lv, 'lv')):
        assert_close(got, T(g[key]), rtol=1e-5, atol=1e-6, what=key)
    assert float(lv.max()) == 10.0 or float(lv.min()) == -10.0          # clamp exercised
    assert_close(oheads.ordinal_probabilities(ordl), T(g['ord_probs']), rtol=1e-5, atol=1e-6, what='ord probs')
    assert_close(oheads.ordinal_severity(ordl), T(g['ord_sev']), rtol=1e-5, atol=1e-6, what='ord severity')
    loss = (cls * T(g['g_cls'])).sum() + (ordl * T(g['g_ord'])).sum() + (mu * T(g['g_mu'])).sum() + (lv * T(g['g_lv'])).sum()
    loss.backward()
    assert_close(x.grad, T(g['dx']), rtol=1e-4, atol=1e-5, what='dx')
    for k, p in P.items():
        assert_close(p.grad, T(g[k + '.grad']), rtol=1e-4, atol=1e-5, what=k)
    counts = {h: sum(v.numel() for k, v in P.items() if k.startswith(h)) for h in ('cls', 'ord', 'unc')}
    assert counts == {'cls': 25220, 'ord': 25091, 'unc': 24962}           # README.md:316


# ----------------------------------------------------------------------------- losses
def test_loss_known_answer():
    o = {'cls_logits': torch.tensor([[2, .5, -1, 0], [.1, .2, .3, .4]]),
         'ordinal_logits': torch.tensor([[1., -1, -2], [.5, .5, -.5]]),
         'mu': torch.tensor([[.5], [2.5]]), 'log_var': torch.tensor([[0.], [-1.]]),
         'kan_severity': torch.tensor([[.3], [2.]])}
    y = torch.tensor([0, 3])
    r = olosses.joint(o, y, y, 4)
    want = {'cls_loss': 0.328758, 'ord_loss': 0.612614, 'unc_loss': -0.017607, 'kan_loss': 0.545000,
            'total_loss': 1.205068}
    g = load_golden('losses.npz')
    for k, v in want.items():
        assert abs(float(r[k]) - v) < 2e-6, k
        assert abs(float(r[k]) - float(g['kat_' + k])) < 1e-6, k


@pytest.mark.parametrize('stage', [1, 2, 3, 4])
def test_losses_match_reference(stage):
    g = load_golden('losses.npz')
    names = ['cls_logits', 'ordinal_logits', 'mu', 'log_var', 'kan_severity']
    o = {k: T(g['in_' + k]).requires_grad_(True) for k in names}
    r = olosses.joint(o, T(g['yc']), T(g['ys']), stage, alpha=T(g['alpha']))
    for k in ('cls_loss', 'ord_loss', 'unc_loss', 'kan_loss', 'total_loss'):
        assert_close(r[k], T(g[f's{stage}_{k}']), rtol=1e-5, atol=1e-6, what=k)
    r['total_loss'].backward()
    for k in names:
        got = o[k].grad if o[k].grad is not None else torch.zeros_like(o[k])
        assert_close(got, T(g[f's{stage}_d_{k}']), rtol=1e-4, atol=1e-7, what='d' + k)


# ----------------------------------------------------------------------------- composed model
@pytest.fixture(scope='module')
def model_sd():
    g = load_golden('model_b2.npz')
    sd = omodel.random_state_dict(int(g['seed']))
    cs = sum(float(sd[k].double().abs().sum()) for k in sorted(sd))
    if abs(cs - float(g['checksum'])) > 1e-6 * float(g['checksum']):
        pytest.skip('torch RNG stream differs from the one the golden file was made with')
    return sd


def test_state_dict_layout(model_sd):
    import os
    from conftest import GOLDEN
    want = [l.split(' ', 1) for l in open(os.path.join(GOLDEN, 'state_dict_keys.txt')).read().splitlines()]
    assert [k for k, _ in want] and set(k for k, _ in want) == set(model_sd)
    for k, shp in want:
        assert str(tuple(model_sd[k].shape)) == shp, k
    n = sum(v.numel() for k, v in model_sd.items() if not k.endswith('knots'))
    assert n == 5706394                                                     # README.md:316
    assert sum(v.numel() for k, v in model_sd.items() if k.startswith('backbone.')) == 5524416


def test_model_forward_matches_reference(model_sd):
    g = load_golden('model_b2.npz')
    images = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(int(g['images_seed'])))
    with torch.no_grad():
        o = omodel.forward(model_sd, images)
        p = omodel.predict(model_sd, images)
    for k in ('cls_logits', 'features', 'ordinal_logits', 'mu', 'log_var', 'kan_severity'):
        assert_close(o[k], T(g['fwd_' + k]), rtol=1e-4, atol=2e-5, what=k)
    assert torch.equal(p['class'], T(g['pred_class']))
    for k in ('class_probs', 'ordinal_probs', 'ordinal_severity', 'uncertainty_mu', 'uncertainty_std', 'kan_severity'):
        assert_close(p[k], T(g['pred_' + k]), rtol=1e-4, atol=2e-5, what=k)


def test_model_loss_and_grads_match_reference(model_sd):
    g = load_golden('model_b2.npz')
    images = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(int(g['images_seed'])))
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not k.endswith('knots') else v)
          for k, v in model_sd.items()}
    o = omodel.forward(sd, images)
    y = torch.tensor([1, 3])
    r = olosses.joint(o, y, y, 4)
    for k in ('cls_loss', 'ord_loss', 'unc_loss', 'kan_loss', 'total_loss'):
        assert_close(r[k], T(g['loss_' + k]), rtol=1e-4, atol=1e-6, what=k)
    r['total_loss'].backward()
    for k in g.files:
        if k.startswith('grad_'):
            assert_close(sd[k[5:]].grad, T(g[k]), rtol=1e-3, atol=1e-6, scale_tol=2e-4, what=k)


# ----------------------------------------------------------------------------- trunk vs independent implementations
@pytest.fixture(scope='module')
def trunk():
    torch.manual_seed(3)
    m = ovit.DeiTTinyOracle().eval()
    # random biases / affine so that every term is exercised (timm init zeroes them)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith('bias') or 'norm' in n:
                p.add_(torch.randn_like(p) * 0.1)
        m.cls_token.normal_(std=0.5)
    x = torch.randn(2, 3, 224, 224)
    with torch.no_grad():
        return m, x, m(x)


def test_trunk_param_count(trunk):
    assert sum(p.numel() for p in trunk[0].parameters()) == 5524416
    assert len(trunk[0].state_dict()) == 150


def test_trunk_functional_equals_module(trunk):
    m, x, y = trunk
    with torch.no_grad():
        assert_close(ovit.forward_functional(m.state_dict(), x), y, rtol=1e-4, atol=1e-5, what='functional')


def test_trunk_vs_torchvision(trunk):
    tv = pytest.importorskip('torchvision.models.vision_transformer')
    m, x, y = trunk
    ref = tv.VisionTransformer(image_size=224, patch_size=16, num_layers=12, num_heads=3, hidden_dim=192, mlp_dim=768)
    ref.heads = torch.nn.Identity()
    missing = ref.load_state_dict(ovit.to_torchvision(m.state_dict()), strict=False)
    assert not missing.unexpected_keys and all(k.startswith('heads') for k in missing.missing_keys)
    with torch.no_grad():
        assert_close(ref.eval()(x), y, rtol=1e-4, atol=1e-5, what='torchvision ViT')


def test_trunk_vs_hf(trunk):
    tr = pytest.importorskip('transformers')
    m, x, y = trunk
    cfg = tr.ViTConfig(hidden_size=192, num_hidden_layers=12, num_attention_heads=3, intermediate_size=768,
                       layer_norm_eps=1e-6, qkv_bias=True, hidden_act='gelu', image_size=224, patch_size=16)
    ref = tr.ViTModel(cfg, add_pooling_layer=False).eval()
    res = ref.load_state_dict(ovit.to_hf_vit(m.state_dict()), strict=False)
    assert not res.unexpected_keys and not res.missing_keys, res
    with torch.no_grad():
        assert_close(ref(pixel_values=x).last_hidden_state[:, 0], y, rtol=1e-4, atol=1e-5, what='HF ViT')
