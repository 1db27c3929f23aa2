"""Randomised oracle-vs-LIVE-reference check (run as a script by tests/test_oracle_vs_live_reference.py, build container only).

The committed fixtures under tests/golden/ pin the oracle on a fixed set of seeded cases.  This script widens that on any machine
that has the reference tree: it imports the reference's own `models/kan.py`, `models/heads.py` and `training/losses.py`
(unmodified, from $ROVIT_REFERENCE or /root/reference), draws random shapes / weights / inputs / stages / focal parameters,
runs reference and oracle side by side on the CPU -- forward AND every gradient -- and prints one JSON object with the worst
deviation per family.  It runs in its own process because the reference's top-level package names (`models`, `training`)
collide with the drop-in aliases other tests install.
"""

import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('ROVIT_REFERENCE', '/root/reference')
sys.path.insert(0, ROOT)
from oracle import heads as oheads  # noqa: E402
from oracle import kan as okan  # noqa: E402
from oracle import losses as olosses  # noqa: E402

sys.path.insert(0, REF)
from models.heads import ClassificationHead, OrdinalHead, UncertaintyHead  # noqa: E402
from models.kan import BSplineBasis, KANLayer, KANSeverityModule  # noqa: E402
from training.losses import JointLoss  # noqa: E402

for m in ('models.kan', 'models.heads', 'training.losses'):
    assert os.path.abspath(sys.modules[m].__file__).startswith(os.path.abspath(REF)), m

torch.set_num_threads(4)
CASES = int(os.environ.get('LIVE_CASES', '12'))


def rel(a, b):
    """max |a - b| / (|b| + scale) with scale = max|b| (+ tiny): an element-wise relative error that is not blown up by
    near-zero entries (loop-vs-einsum reordering alone gives 1.9e-3 relative on near-zero outputs)."""
    a, b = a.detach().double(), b.detach().double()
    if b.numel() == 0:
        return 0.0
    return float(((a - b).abs() / (b.abs() + b.abs().max() + 1e-12)).max())


def check_basis(rng):
    knots = torch.linspace(-1, 1, 11)
    worst = 0.0
    dead_zone_ok = True
    for _ in range(CASES):
        b, d = int(torch.randint(1, 40, (1,), generator=rng)), int(torch.randint(1, 30, (1,), generator=rng))
        t = torch.tanh(torch.randn(b, d, generator=rng) * 1.5)
        # plant values on / next to knots and the jump at 0.4 (= knots[7]), and outside the clamp range
        flat = t.view(-1)
        special = torch.cat([knots, knots - 1e-6, knots + 1e-6, torch.tensor([0.399999, 0.4, 0.400001, -1.5, 1.5])])
        n = min(flat.numel(), special.numel())
        flat[:n] = special[torch.randperm(special.numel(), generator=rng)[:n]]
        want = BSplineBasis.compute_basis(t, knots, 3)
        dead = torch.clamp(t, knots[0], knots[-1]) >= knots[7]     # the reference's missing degree-0 seeds (kan.py:12-13, 23-25)
        alive = (t > knots[0] + 1e-3) & ~dead
        for fn in (BSplineBasis.compute_basis, okan.basis_literal, okan.basis_closed_form):
            got = fn(t, knots)
            worst = max(worst, float((got - want).abs().max()))
            dead_zone_ok &= bool((got[dead] == 0).all()) and bool((got[alive].sum(-1) > 0).all())
    return {'max_abs': worst, 'dead_zone_exactly_zero_from_0.4': dead_zone_ok}


def check_kan_layer(rng):
    worst = {'y': 0.0, 'dx': 0.0, 'dsw': 0.0, 'dlw': 0.0, 'dlb': 0.0}
    for c in range(CASES):
        n_in, n_out, b = (int(torch.randint(lo, hi, (1,), generator=rng)) for lo, hi in ((1, 20), (1, 7), (1, 12)))
        torch.manual_seed(1000 + c)
        layer = KANLayer(n_in, n_out)
        x = (torch.randn(b, n_in, generator=rng) * 1.3).requires_grad_(True)
        gy = torch.randn(b, n_out, generator=rng)
        y = layer(x)
        y.backward(gy)
        for loop in (False, True):
            xo = x.detach().clone().requires_grad_(True)
            ps = [p.detach().clone().requires_grad_(True) for p in (layer.spline_weights, layer.linear.weight, layer.linear.bias)]
            fn = okan.layer_forward_loop if loop else okan.layer_forward
            yo = fn(xo, *ps, layer.knots)
            yo.backward(gy)
            for k, a, r in (('y', yo, y), ('dx', xo.grad, x.grad), ('dsw', ps[0].grad, layer.spline_weights.grad),
                            ('dlw', ps[1].grad, layer.linear.weight.grad), ('dlb', ps[2].grad, layer.linear.bias.grad)):
                worst[k] = max(worst[k], rel(a, r))
    return worst


def check_kan_module(rng):
    worst = {'y': 0.0, 'dx': 0.0, 'trajectory': 0.0}
    for c in range(max(CASES // 2, 3)):
        depth = int(torch.randint(1, 4, (1,), generator=rng))
        dims = [int(torch.randint(2, 14, (1,), generator=rng)) for _ in range(depth)] + [1]
        torch.manual_seed(2000 + c)
        mod = KANSeverityModule(dims).eval()
        x = torch.randn(5, dims[0], generator=rng, requires_grad=True)
        y = mod(x)
        y.sum().backward()
        layers = [(l.spline_weights.detach(), l.linear.weight.detach(), l.linear.bias.detach()) for l in mod.kan_layers]
        xo = x.detach().clone().requires_grad_(True)
        traj = okan.severity_forward(xo, layers, mod.kan_layers[0].knots, return_trajectory=True)
        traj[-1].sum().backward()
        worst['y'] = max(worst['y'], rel(traj[-1], y))
        worst['dx'] = max(worst['dx'], rel(xo.grad, x.grad))
        assert float(y.detach().min()) >= 0.0 and float(y.detach().max()) <= 3.0
        with torch.no_grad():
            for a, r in zip(traj, mod.get_activation_trajectory(x.detach())):
                worst['trajectory'] = max(worst['trajectory'], rel(a, r))
    return worst


def check_heads(rng):
    worst = 0.0
    for c in range(CASES):
        e, h, k, b = (int(torch.randint(lo, hi, (1,), generator=rng)) for lo, hi in ((4, 40), (2, 24), (2, 7), (1, 9)))
        torch.manual_seed(3000 + c)
        ch, oh, uh = ClassificationHead(e, h, k).eval(), OrdinalHead(e, h, k).eval(), UncertaintyHead(e, h).eval()
        with torch.no_grad():
            uh.fc_logvar.weight.mul_(40.0)          # drive some log-variances into the clamp at +-10 (heads.py:99)
        x = torch.randn(b, e, generator=rng, requires_grad=True)
        outs_ref = (ch(x), oh(x), *uh(x))
        xo = x.detach().clone().requires_grad_(True)
        w = lambda lin: (lin.weight.detach(), lin.bias.detach())
        mu, lv = oheads.uncertainty_forward(xo, *w(uh.fc1), *w(uh.fc_mu), *w(uh.fc_logvar))
        outs = (oheads.classification_forward(xo, *w(ch.fc1), *w(ch.fc2)), oheads.ordinal_forward(xo, *w(oh.fc1), *w(oh.fc2)), mu, lv)
        g = [torch.randn(o.shape, generator=rng) for o in outs_ref]
        torch.autograd.backward(outs_ref, g)
        torch.autograd.backward(outs, g)
        worst = max([worst, rel(xo.grad, x.grad)] + [rel(a, r) for a, r in zip(outs, outs_ref)])
        with torch.no_grad():
            worst = max(worst, rel(oheads.ordinal_probabilities(outs[1]), oh.predict_probabilities(x)),
                        rel(oheads.ordinal_severity(outs[1]), oh.predict_severity(x)))
    return {'max_rel': worst}


def check_losses(rng):
    worst, worst_grad = 0.0, 0.0
    for c in range(CASES * 2):
        b, k = int(torch.randint(1, 17, (1,), generator=rng)), 4
        stage = 1 + c % 4
        gamma = (0.5, 1.0, 2.0, 3.0)[int(torch.randint(0, 4, (1,), generator=rng))]
        alpha = None if c % 3 == 0 else torch.rand(k, generator=rng) + 0.25
        lam = [float(v) for v in torch.rand(3, generator=rng) * 2]
        mk = lambda *s: (torch.randn(*s, generator=rng) * 1.5).requires_grad_(True)
        outs = {'cls_logits': mk(b, k), 'ordinal_logits': mk(b, k - 1), 'mu': mk(b, 1), 'log_var': mk(b, 1),
                'kan_severity': (torch.rand(b, 1, generator=rng) * 3).requires_grad_(True)}
        if c % 5 == 4:                      # a head the model gated off: the reference tests `is not None` (losses.py:153-175)
            outs['kan_severity'] = None
        y = torch.randint(0, k, (b,), generator=rng)
        sev = torch.randint(0, k, (b,), generator=rng)
        ref = JointLoss(lam[0], lam[1], lam[2], focal_gamma=gamma, focal_alpha=alpha)(outs, y, sev, stage)
        ref['total_loss'].backward()
        outs_o = {n: (None if v is None else v.detach().clone().requires_grad_(True)) for n, v in outs.items()}
        got = olosses.joint(outs_o, y, sev, stage, lam[0], lam[1], lam[2], gamma, alpha)
        got['total_loss'].backward()
        assert set(got) == set(ref)
        for n in ref:
            worst = max(worst, abs(float(got[n].detach()) - float(ref[n].detach())) / (abs(float(ref[n].detach())) + 1e-3))
        for n, v in outs.items():
            if v is None:
                continue
            if v.grad is None:
                assert outs_o[n].grad is None or float(outs_o[n].grad.abs().max()) == 0.0, n
            else:
                worst_grad = max(worst_grad, rel(outs_o[n].grad, v.grad))
    return {'max_rel_loss': worst, 'max_rel_grad': worst_grad}


def check_model(rng):
    """The composed RoViTKAN of the reference (models/rovit_kan.py, unmodified; the absent third-party timm replaced by the oracle's
    restatement of deit_tiny_patch16_224) against oracle.model.forward / predict at every curriculum stage: same keys, the same
    entries gated to None, same numbers."""
    import types
    from oracle import model as omodel
    from oracle import vit as ovit
    shim = types.ModuleType('timm')
    shim.create_model = ovit.create_model
    sys.modules['timm'] = shim
    from models.rovit_kan import RoViTKAN
    assert os.path.abspath(sys.modules['models.rovit_kan'].__file__).startswith(os.path.abspath(REF))
    sd = omodel.random_state_dict(int(torch.randint(0, 1000, (1,), generator=rng)))
    with torch.no_grad():
        for k, v in sd.items():            # timm initialises every bias to zero: move them so that they count
            if k.endswith('bias'):
                v.add_(torch.randn(v.shape, generator=rng) * 0.05)
    ref = RoViTKAN(pretrained=False, dropout=0.0)
    ref.load_state_dict(sd, strict=True)
    ref.eval()
    x = torch.randn(2, 3, 224, 224, generator=rng)
    worst, gating_ok, keys_ok = 0.0, True, True
    for stage in (1, 2, 3, 4):
        ref.curriculum_stage = stage
        with torch.no_grad():
            a, b = ref(x), omodel.forward(sd, x, stage=stage)
            pa, pb = ref.predict(x), omodel.predict(sd, x, stage=stage)
        keys_ok &= set(a) == set(b) and set(pa) == set(pb)
        for d_ref, d_or in ((a, b), (pa, pb)):
            for k in d_ref:
                gating_ok &= (d_ref[k] is None) == (d_or[k] is None)
                if d_ref[k] is None:
                    continue
                if d_ref[k].dtype == torch.int64:
                    keys_ok &= bool(torch.equal(d_ref[k], d_or[k]))
                else:
                    worst = max(worst, rel(d_or[k], d_ref[k]))
    return {'max_rel': worst, 'stage_gating_matches': gating_ok, 'keys_and_classes_match': keys_ok}


def main():
    rng = torch.Generator().manual_seed(int(os.environ.get('LIVE_SEED', '0')))
    out = {'reference': REF, 'cases': CASES, 'basis': check_basis(rng), 'kan_layer': check_kan_layer(rng),
           'kan_module': check_kan_module(rng), 'heads': check_heads(rng), 'losses': check_losses(rng), 'model': check_model(rng)}
    print(json.dumps(out))


if __name__ == '__main__':
    main()
