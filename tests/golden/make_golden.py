"""Generate the golden vectors under tests/golden/ from the REFERENCE itself.

Run in the build container only (needs /root/reference; the GPU box has none):

    python tests/golden/make_golden.py

It imports the reference's own `models/kan.py`, `models/heads.py`,
`training/losses.py` unmodified, and `models/rovit_kan.py` with the absent
third-party `timm` replaced by the oracle's restatement of
`deit_tiny_patch16_224` (oracle/vit.py).  Everything it stores is an output of
reference code on seeded inputs; tests compare the oracle and the CUDA path to
these files.  Large weights are not stored: they are re-derived from the seed
and guarded by a checksum.
"""

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('ROVIT_REFERENCE', '/root/reference')
sys.path.insert(0, ROOT)

from oracle import vit as ovit  # noqa: E402

# timm is not installed: give the reference the oracle's restatement instead.
shim = types.ModuleType('timm')
shim.create_model = ovit.create_model
sys.modules['timm'] = shim
sys.path.insert(0, REF)

from models.kan import BSplineBasis, KANLayer, KANSeverityModule  # noqa: E402
from models.heads import ClassificationHead, OrdinalHead, UncertaintyHead  # noqa: E402
from models.rovit_kan import RoViTKAN  # noqa: E402
from training.losses import JointLoss  # noqa: E402

torch.set_num_threads(8)


def npy(t):
    return t.detach().cpu().numpy()


def gen_basis():
    knots = torch.linspace(-1, 1, 11)
    grid = torch.cat([torch.linspace(-1, 1, 21), knots, knots - 1e-6, knots + 1e-6,
                      torch.tensor([-0.95, -0.05, 0.05, 0.35, 0.399999, 0.4, 0.400001, 0.97])])
    g = torch.Generator().manual_seed(1)
    rnd = torch.tanh(torch.randn(4096, generator=g) * 1.5)
    t = torch.cat([grid, rnd])[None, :]
    basis = BSplineBasis.compute_basis(t, knots, 3)
    np.savez_compressed(os.path.join(HERE, 'kan_basis.npz'), t=npy(t), knots=npy(knots), basis=npy(basis))


def gen_kan():
    out = {}
    for tag, (n_in, n_out, bsz) in {'l0': (192, 64, 16), 'l1': (64, 16, 32), 'l2': (16, 1, 64),
                                    'odd': (10, 3, 7)}.items():
        torch.manual_seed(10)
        layer = KANLayer(n_in, n_out)
        x = (torch.randn(bsz, n_in) * 1.2).requires_grad_(True)
        gy = torch.randn(bsz, n_out)
        y = layer(x)
        y.backward(gy)
        out.update({f'{tag}_x': npy(x), f'{tag}_gy': npy(gy), f'{tag}_y': npy(y), f'{tag}_dx': npy(x.grad),
                    f'{tag}_sw': npy(layer.spline_weights), f'{tag}_lw': npy(layer.linear.weight),
                    f'{tag}_lb': npy(layer.linear.bias), f'{tag}_knots': npy(layer.knots),
                    f'{tag}_dsw': npy(layer.spline_weights.grad), f'{tag}_dlw': npy(layer.linear.weight.grad),
                    f'{tag}_dlb': npy(layer.linear.bias.grad)})
    torch.manual_seed(11)
    mod = KANSeverityModule([192, 64, 16, 1])
    x = torch.randn(8, 192, requires_grad=True)
    gy = torch.randn(8, 1)
    y = mod(x)
    y.backward(gy)
    out.update({'mod_x': npy(x), 'mod_gy': npy(gy), 'mod_y': npy(y), 'mod_dx': npy(x.grad)})
    for i, l in enumerate(mod.kan_layers):
        out.update({f'mod_sw{i}': npy(l.spline_weights), f'mod_lw{i}': npy(l.linear.weight),
                    f'mod_lb{i}': npy(l.linear.bias), f'mod_dsw{i}': npy(l.spline_weights.grad),
                    f'mod_dlw{i}': npy(l.linear.weight.grad), f'mod_dlb{i}': npy(l.linear.bias.grad)})
    traj = mod.get_activation_trajectory(x.detach())
    for i, a in enumerate(traj):
        out[f'mod_traj{i}'] = npy(a)
    np.savez_compressed(os.path.join(HERE, 'kan_layers.npz'), **out)


def gen_heads():
    torch.manual_seed(20)
    x = torch.randn(12, 192, requires_grad=True)
    out = {'x': npy(x)}
    ch = ClassificationHead(192, 128, 4, dropout=0.0)
    oh = OrdinalHead(192, 128, 4, dropout=0.0)
    uh = UncertaintyHead(192, 128, dropout=0.0)
    for m in (ch, oh, uh):
        m.train()
    g_cls, g_ord, g_mu, g_lv = torch.randn(12, 4), torch.randn(12, 3), torch.randn(12, 1), torch.randn(12, 1)
    cls = ch(x)
    ordl = oh(x)
    # push a few log-variances outside the clamp so that its zero-gradient region is exercised
    with torch.no_grad():
        uh.fc_logvar.weight.mul_(60.0)
    mu, lv = uh(x)
    (cls * g_cls).sum().backward(retain_graph=True)
    (ordl * g_ord).sum().backward(retain_graph=True)
    ((mu * g_mu).sum() + (lv * g_lv).sum()).backward()
    out.update({'cls': npy(cls), 'ord': npy(ordl), 'mu': npy(mu), 'lv': npy(lv), 'dx': npy(x.grad),
                'g_cls': npy(g_cls), 'g_ord': npy(g_ord), 'g_mu': npy(g_mu), 'g_lv': npy(g_lv),
                'ord_probs': npy(oh.predict_probabilities(x)), 'ord_sev': npy(oh.predict_severity(x))})
    for name, m in (('cls', ch), ('ord', oh), ('unc', uh)):
        for k, p in m.named_parameters():
            out[f'{name}.{k}'] = npy(p)
            out[f'{name}.{k}.grad'] = npy(p.grad)
    np.savez_compressed(os.path.join(HERE, 'heads.npz'), **out)


def gen_losses():
    out = {}
    # RNG-free known-answer case from SURVEY.md section 4 (1b)
    kat = {'cls_logits': torch.tensor([[2, .5, -1, 0], [.1, .2, .3, .4]]),
           'ordinal_logits': torch.tensor([[1., -1, -2], [.5, .5, -.5]]),
           'mu': torch.tensor([[.5], [2.5]]), 'log_var': torch.tensor([[0.], [-1.]]),
           'kan_severity': torch.tensor([[.3], [2.]])}
    y = torch.tensor([0, 3])
    res = JointLoss(focal_alpha=None)(kat, y, y, 4)
    for k, v in res.items():
        out['kat_' + k] = npy(v)
    # seeded case with gradients, per-class alpha, every stage
    torch.manual_seed(30)
    bsz = 37
    o = {'cls_logits': torch.randn(bsz, 4) * 2, 'ordinal_logits': torch.randn(bsz, 3) * 2,
         'mu': torch.randn(bsz, 1) + 1.5, 'log_var': torch.randn(bsz, 1), 'kan_severity': torch.rand(bsz, 1) * 3}
    yc = torch.randint(0, 4, (bsz,))
    ys = torch.randint(0, 4, (bsz,))
    alpha = torch.tensor([0.7, 1.3, 1.0, 2.1])
    for k, v in o.items():
        out['in_' + k] = npy(v)
    out['yc'], out['ys'], out['alpha'] = npy(yc), npy(ys), npy(alpha)
    for stage in (1, 2, 3, 4):
        oo = {k: v.clone().requires_grad_(True) for k, v in o.items()}
        res = JointLoss(focal_alpha=alpha)(oo, yc, ys, stage)
        res['total_loss'].backward()
        for k, v in res.items():
            out[f's{stage}_{k}'] = npy(v)
        for k, v in oo.items():
            out[f's{stage}_d_{k}'] = npy(v.grad) if v.grad is not None else np.zeros(v.shape, np.float32)
    np.savez_compressed(os.path.join(HERE, 'losses.npz'), **out)


def weights_checksum(sd):
    acc = 0.0
    for k in sorted(sd):
        acc += float(sd[k].double().abs().sum())
    return acc


def gen_model():
    """Full reference RoViTKAN (reference composition + reference heads/KAN; trunk =
    oracle restatement of timm) on a 2-image batch.  Weights are re-derived from
    the seed by tests (construction order below is what they replay)."""
    seed = 0
    torch.manual_seed(seed)
    model = RoViTKAN(pretrained=False, dropout=0.0)
    model.eval()
    sd = model.state_dict()
    g = torch.Generator().manual_seed(1234)
    images = torch.randn(2, 3, 224, 224, generator=g)
    out = {'seed': np.int64(seed), 'checksum': np.float64(weights_checksum(sd)), 'images_seed': np.int64(1234)}
    with torch.no_grad():
        o = model(images)
        p = model.predict(images)
    for k, v in o.items():
        out['fwd_' + k] = npy(v)
    for k, v in p.items():
        out['pred_' + k] = npy(v)
    # a training-style pass: joint loss and a handful of gradients
    model.train()
    yc = torch.tensor([1, 3])
    o = model(images)
    res = JointLoss(focal_alpha=None)(o, yc, yc, 4)
    res['total_loss'].backward()
    for k, v in res.items():
        out['loss_' + k] = npy(v)
    named = dict(model.named_parameters())
    for k in ['backbone.model.cls_token', 'backbone.model.pos_embed', 'backbone.model.patch_embed.proj.bias',
              'backbone.model.blocks.0.norm1.weight', 'backbone.model.blocks.0.attn.qkv.bias',
              'backbone.model.blocks.5.attn.proj.weight', 'backbone.model.blocks.11.mlp.fc1.bias',
              'backbone.model.blocks.11.mlp.fc2.weight', 'backbone.model.norm.weight',
              'classification_head.fc1.weight', 'ordinal_head.fc2.weight', 'uncertainty_head.fc_logvar.weight',
              'kan_module.kan_layers.0.spline_weights', 'kan_module.kan_layers.2.linear.weight']:
        out['grad_' + k] = npy(named[k].grad)
    # same pass at curriculum stage 3 (no KAN branch): the KAN's discontinuity at tanh(x)=0.4 makes stage-4
    # gradients hypersensitive to bf16-sized feature noise, stage-3 gradients are the well-conditioned check
    model.zero_grad()
    model.curriculum_stage = 3
    o3 = model(images)
    res3 = JointLoss(focal_alpha=None)(o3, yc, yc, 3)
    res3['total_loss'].backward()
    for k, v in res3.items():
        out['loss3_' + k] = npy(v)
    for k in [n for n in out if n.startswith('grad_')]:
        g3 = named[k[5:]].grad
        if g3 is not None and not k[5:].startswith('kan_module'):
            out['grad3_' + k[5:]] = npy(g3)
    model.curriculum_stage = 4
    out['param_count'] = np.int64(sum(p.numel() for p in model.parameters()))
    out['param_counts'] = np.array([model.count_parameters()[k] for k in
                                    ['backbone', 'classification_head', 'ordinal_head', 'uncertainty_head',
                                     'kan_module', 'total']], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, 'model_b2.npz'), **out)
    names = list(sd.keys())
    with open(os.path.join(HERE, 'state_dict_keys.txt'), 'w') as f:
        for k in names:
            f.write(f'{k} {tuple(sd[k].shape)}\n')


if __name__ == '__main__':
    gen_basis()
    gen_kan()
    gen_heads()
    gen_losses()
    gen_model()
    print('golden vectors written to', HERE)
