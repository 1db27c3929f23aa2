"""End-to-end parity at the BASELINE.json sizes (VERDICT r1 item 1a / SURVEY 8c "parity procedure").

  cfg 2  inference, batch 1024: every output against the fp32 oracle run on the same device (TF32 off); class argmax and
         BOTH ordinal decodes (argmax of predict_probabilities, count of positive cumulative logits -- the reference defines
         no discrete ordinal prediction) must agree on every row whose reference margin exceeds a stated bound;
         the agreement rate over ALL rows and the worst margin of any mismatch are printed.
  cfg 4  stage-4 train step, batch 256: five losses and all 173 parameter gradients.
  cfg 3  KANSeverityModule([192,64,1]), batch 65536, forward + backward against the oracle's autograd.

Second yard-stick: our deviation from the fp32 oracle must not exceed the deviation of the oracle trunk itself
run under torch.autocast(bfloat16) on the same box (YARDSTICK_SLACK x).

The reference's KAN basis is discontinuous at tanh(x) = 0.4.  A sample is "flip-free" when, for every KAN layer,
our layer inputs and the oracle's fall on the same side of that jump; `kan_severity` (and the KAN term of the train-step
loss) is compared tightly on the flip-free samples, and the flip rate is reported and bounded.
"""

import pytest
import torch

from conftest import assert_close
from oracle import kan as okan
from oracle import losses as olosses
from oracle import model as omodel
from oracle import vit as ovit

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from rovitkan_b200.models import RoViTKAN
    from rovitkan_b200.models.kan import KANSeverityModule
    from rovitkan_b200.training.losses import JointLoss

DEV = 'cuda'
# stated bf16-trunk tolerances = 2x the worst deviation measured on B200 at these sizes (round 2, batch 1024: rel-L2
# 1.6e-3 .. 7.4e-3 per output tensor, worst element 0.85e-2 of |b| + max|b|; the oracle trunk under autocast(bf16) on the
# same box: 3.1e-3 .. 1.7e-2; flip-free kan_severity max 2.8e-2; 38 % of the samples see a basis flip in one of the three
# KAN layers at batch 1024, 43 % at batch 256)
OUT_RTOL, OUT_STOL = 2e-2, 2e-2          # |a-b| <= OUT_RTOL*|b| + OUT_STOL*max|b|   (features, logits, mu, log_var)
OUT_REL_L2 = 1.5e-2                      # relative L2 per output tensor
KAN_CLEAN_ATOL = 6e-2                    # kan_severity on flip-free samples ([0,3] range)
KAN_FLIP_RATE_MAX = 0.6                  # share of samples with at least one basis flip (reported; bounded loosely)
GRAD_REL_L2_TRUNK = 6e-2                 # per trunk gradient tensor, stage-3 loss at batch 256 (measured: median 4.4e-3, worst
                                         # 3.4e-2 = patch_embed.proj.weight, the end of the 12-block chain)
GRAD_REL_L2_HEADS = 5e-2                 # per head gradient tensor (measured worst 2.4e-2)
YARDSTICK_SLACK = 1.0                    # our error must not exceed the bf16-autocast reference's error at all


def rel_l2(a, b):
    a, b = a.detach().double().flatten(), b.detach().double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _weights(seed):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    sd = omodel.random_state_dict(seed)
    torch.manual_seed(seed + 1)
    with torch.no_grad():          # non-trivial biases / affines; wider heads so the decisions are not all near-ties
        for k, v in sd.items():
            if k.startswith('backbone') and (k.endswith('bias') or 'norm' in k):
                v.add_(torch.randn_like(v) * 0.05)
    return sd


def _model(sd, train=False):
    m = RoViTKAN(pretrained=False, dropout=0.0)
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV)
    return m.train() if train else m.eval()


def _oracle_forward(sdd, images, chunk=128, autocast=False):
    outs = []
    with torch.no_grad():
        for i in range(0, images.shape[0], chunk):
            x = images[i:i + chunk]
            if autocast:
                with torch.autocast('cuda', dtype=torch.bfloat16):
                    f = ovit.forward_functional(sdd, x, prefix='backbone.model.').float()
                outs.append(omodel.heads_forward(sdd, f, 4))
            else:
                outs.append(omodel.forward(sdd, x))
    return {k: torch.cat([o[k] for o in outs]) for k in outs[0]}


def _kan_inputs_oracle(sdd, features):
    """Inputs of every KAN layer in the oracle (kan.py:138-149: ReLU between layers)."""
    xs, x = [], features
    layers = omodel.kan_layers_from(sdd)
    knots = sdd['kan_module.kan_layers.0.knots']
    for i, (sw, lw, lb) in enumerate(layers):
        xs.append(x)
        if i + 1 < len(layers):
            x = torch.relu(okan.layer_forward(x, sw, lw, lb, knots))
    return xs


def _flip_free(model, sdd, feats_ours, feats_ref):
    """True per sample when no KAN-layer input crosses the tanh(x) = knots[7] jump between our run and the oracle's."""
    knots = sdd['kan_module.kan_layers.0.knots']
    edge = float(knots[7])
    with torch.no_grad():
        ours = model.kan_module.get_activation_trajectory(feats_ours)[:-1]
        ref = _kan_inputs_oracle(sdd, feats_ref)
    clean = torch.ones(feats_ours.shape[0], dtype=torch.bool, device=feats_ours.device)
    for a, b in zip(ours, ref):
        clean &= ((torch.tanh(a) >= edge) == (torch.tanh(b) >= edge)).all(dim=1)
    return clean


def _decisions(o):
    cls = o['cls_logits']
    top2 = cls.topk(2, dim=1).values
    c = torch.sigmoid(o['ordinal_logits'])
    probs = torch.cat([c[:, :1], c[:, 1:] - c[:, :-1], 1.0 - c[:, -1:]], dim=1)       # heads.py:45-67
    ptop2 = probs.topk(2, dim=1).values
    return {'class': (cls.argmax(1), top2[:, 0] - top2[:, 1]),
            'ordinal argmax(predict_probabilities)': (probs.argmax(1), ptop2[:, 0] - ptop2[:, 1]),
            'ordinal count(logit>0)': ((o['ordinal_logits'] > 0).sum(1), o['ordinal_logits'].abs().min(dim=1).values)}


def test_inference_batch_1024_against_fp32_oracle_and_bf16_yardstick():
    sd = _weights(0)
    sdd = {k: v.to(DEV) for k, v in sd.items()}
    images = torch.randn(1024, 3, 224, 224, generator=torch.Generator().manual_seed(0)).to(DEV)
    m = _model(sd)
    with torch.no_grad():
        o = m(images)
    ref = _oracle_forward(sdd, images)
    yard = _oracle_forward(sdd, images, autocast=True)
    print()
    for k in ('features', 'cls_logits', 'ordinal_logits', 'mu', 'log_var'):
        e, ey = rel_l2(o[k], ref[k]), rel_l2(yard[k], ref[k])
        worst = float(((o[k] - ref[k]).abs() / (OUT_RTOL * ref[k].abs() + OUT_STOL * ref[k].abs().max())).max())
        print(f'  {k:15s} rel-L2 ours {e:.2e} | oracle under autocast(bf16) {ey:.2e} | worst elementwise err / tolerance {worst:.2f}')
        assert e <= OUT_REL_L2, (k, e)
        assert e <= YARDSTICK_SLACK * ey, f'{k}: our error {e:.2e} exceeds {YARDSTICK_SLACK}x the bf16-autocast reference error {ey:.2e}'
        assert_close(o[k], ref[k], rtol=OUT_RTOL, atol=0, scale_tol=OUT_STOL, what=k)
    # decisions
    dec_o, dec_r = _decisions(o), _decisions(ref)
    for name in dec_r:
        (po, _), (pr, margin) = dec_o[name], dec_r[name]
        scale = float(ref['cls_logits'].abs().max() if name == 'class' else (1.0 if 'argmax' in name else ref['ordinal_logits'].abs().max()))
        bound = 2 * OUT_STOL * scale
        mism = po != pr
        rate = 1.0 - float(mism.float().mean())
        worst = float(margin[mism].max()) if bool(mism.any()) else 0.0
        decided = margin > bound
        print(f'  {name:38s} agreement {rate:.4f} over 1024 rows; {int(decided.sum())} rows decided beyond {bound:.3g}; '
              f'worst reference margin of a mismatch {worst:.3g}')
        assert torch.equal(po[decided], pr[decided]), name
        assert rate >= 0.97, (name, rate)
    # KAN severity: tight on flip-free samples, flips counted
    clean = _flip_free(m, sdd, o['features'], ref['features'])
    dev = (o['kan_severity'] - ref['kan_severity']).abs().flatten()
    dev_y = (yard['kan_severity'] - ref['kan_severity']).abs().flatten()
    flip_rate = 1.0 - float(clean.float().mean())
    print(f'  kan_severity: flip-free samples {int(clean.sum())}/1024; |dev| on them max {float(dev[clean].max()):.2e} mean '
          f'{float(dev[clean].mean()):.2e}; on flipped samples max {float(dev[~clean].max()) if bool((~clean).any()) else 0.0:.2e}; '
          f'all-sample mean ours {float(dev.mean()):.2e} vs bf16-autocast reference {float(dev_y.mean()):.2e}')
    assert flip_rate <= KAN_FLIP_RATE_MAX
    assert float(dev[clean].max()) <= KAN_CLEAN_ATOL
    assert float(dev.mean()) <= YARDSTICK_SLACK * float(dev_y.mean()) + 1e-3
    assert float(dev.max()) <= 0.35
    # given the oracle's own features the fp32 tail is exact (1e-3) at this batch too, decisions bit-exact
    with torch.no_grad():
        t = m.kan_module(ref['features'])
        cls = m.classification_head(ref['features'])
        ordl = m.ordinal_head(ref['features'])
    assert_close(t, ref['kan_severity'], rtol=1e-3, atol=1e-5, what='kan_severity given reference features')
    assert torch.equal(cls.argmax(1), ref['cls_logits'].argmax(1))
    assert torch.equal((ordl > 0).sum(1), (ref['ordinal_logits'] > 0).sum(1))


def test_train_step_batch_256_losses_and_all_gradients():
    """BASELINE configs[3] at full size.  Two backward passes per implementation:
      (i)  the stage-3 loss (focal + ordinal + uncertainty): well conditioned -> absolute bounds on all 164 gradients it reaches;
      (ii) the full stage-4 loss.  The reference's KAN makes d(loss)/d(features) hypersensitive: with the ORACLE alone, relative
           feature noise of 1e-4 already moves the KAN gradients by 0.5 %, 2.5e-3 (the size of any bf16 trunk's error) by 10 %
           (measured; cascaded cubic segments of width 0.2 plus the jump at tanh(x) = 0.4).  An absolute bound would only
           measure that conditioning, so stage 4 is judged by the survey's second yard-stick: the reference's own trunk under
           torch.autocast(bfloat16) on this device, against the same fp32 oracle."""
    batch = 256
    sd = _weights(4)
    g = torch.Generator().manual_seed(4)
    images = torch.randn(batch, 3, 224, 224, generator=g).to(DEV)
    yc = torch.randint(0, 4, (batch,), generator=g).to(DEV)
    alpha = torch.tensor([0.7, 1.1, 0.9, 1.3], device=DEV)

    def oracle_run(stage, autocast):
        sdd = {k: (v.to(DEV).requires_grad_(True) if not k.endswith('knots') else v.to(DEV)) for k, v in sd.items()}
        if autocast:
            with torch.autocast('cuda', dtype=torch.bfloat16):
                f = ovit.forward_functional(sdd, images, prefix='backbone.model.')
            out = omodel.heads_forward(sdd, f.float(), 4)
        else:
            out = omodel.forward(sdd, images)
        losses = olosses.joint(out, yc, yc, stage, alpha=alpha)
        losses['total_loss'].backward()
        return {k: v.detach() for k, v in losses.items()}, {k: v.grad for k, v in sdd.items() if v.requires_grad and v.grad is not None}

    def ours_run(stage):
        m = _model(sd, train=True)
        losses = JointLoss(focal_alpha=alpha)(m(images), yc, yc, stage)
        losses['total_loss'].backward()
        return {k: v.detach() for k, v in losses.items()}, {k: p.grad for k, p in m.named_parameters() if p.grad is not None}

    print()
    # ---- (i) stage-3 loss
    l_ref, g_ref = oracle_run(3, False)
    l_our, g_our = ours_run(3)
    for k in ('cls_loss', 'ord_loss', 'unc_loss', 'total_loss'):
        print(f'  stage 3 {k}: ours {float(l_our[k]):.6f} oracle {float(l_ref[k]):.6f}')
        assert_close(l_our[k], l_ref[k], rtol=1e-2, atol=1e-3, what=k)
    assert set(g_our) == set(g_ref) and len(g_our) == 173 - 9
    errs = {k: rel_l2(g_our[k], g_ref[k]) for k in g_ref}
    trunk = {k: v for k, v in errs.items() if k.startswith('backbone')}
    heads = {k: v for k, v in errs.items() if not k.startswith('backbone')}
    print('  stage 3 trunk gradients (150): worst', max((v, k) for k, v in trunk.items()), 'median', sorted(trunk.values())[75])
    print('  stage 3 head gradients (14): worst', max((v, k) for k, v in heads.items()))
    bad = {k: v for k, v in trunk.items() if not v <= GRAD_REL_L2_TRUNK}
    bad.update({k: v for k, v in heads.items() if not v <= GRAD_REL_L2_HEADS})
    assert not bad, bad
    # ---- (ii) stage-4 loss against the bf16-autocast yard-stick
    l_ref, g_ref = oracle_run(4, False)
    l_yard, g_yard = oracle_run(4, True)
    l_our, g_our = ours_run(4)
    assert len(g_our) == 173
    for k in ('cls_loss', 'ord_loss', 'unc_loss', 'kan_loss', 'total_loss'):
        print(f'  stage 4 {k}: ours {float(l_our[k]):.6f} oracle {float(l_ref[k]):.6f} oracle under autocast(bf16) {float(l_yard[k]):.6f}')
        assert abs(float(l_our[k]) - float(l_ref[k])) <= max(1.5 * abs(float(l_yard[k]) - float(l_ref[k])), 1e-2 * abs(float(l_ref[k])) + 1e-3), k
    e_our = {k: rel_l2(g_our[k], g_ref[k]) for k in g_ref}
    e_yard = {k: rel_l2(g_yard[k], g_ref[k]) for k in g_ref}
    ratio = sorted(e_our[k] / max(e_yard[k], 1e-12) for k in g_ref)
    print('  stage 4 gradient rel-L2 vs fp32 oracle: ours median', sorted(e_our.values())[86], 'worst', max((v, k) for k, v in e_our.items()))
    print('                  oracle under autocast(bf16): median', sorted(e_yard.values())[86], 'worst', max((v, k) for k, v in e_yard.items()))
    print(f'  ours / yard-stick per tensor: median {ratio[86]:.2f}, 90th percentile {ratio[155]:.2f}, max {ratio[-1]:.2f}')
    assert ratio[86] <= 1.0, 'half of the gradient tensors are further from the fp32 reference than the reference under bf16 autocast'
    assert ratio[155] <= 1.5 and all(torch.isfinite(v).all() for v in g_our.values())


def test_kan_microbench_config_batch_65536_forward_backward():
    """BASELINE configs[2]: KANSeverityModule([192,64,1]) at batch 65536, fp32 tolerance 1e-3 (the tensor-core kernels split
    their operands hi+lo).  x is identical on both sides, so interval decisions can only differ where tanhf and torch.tanh
    disagree in the last ulp exactly at a knot, and -- through the stack -- where a hidden activation of the tensor-core first
    layer lands within its 1e-6 rounding of the tanh(x) = 0.4 jump or of the ReLU gate (measured: 15 of 65536 samples); at
    most OUTLIERS samples may miss the tolerance."""
    OUTLIERS = 48
    batch = 65536
    torch.manual_seed(0)
    mod = KANSeverityModule([192, 64, 1]).to(DEV)
    x = torch.randn(batch, 192, generator=torch.Generator().manual_seed(0)).to(DEV).requires_grad_(True)
    gy = torch.randn(batch, 1, generator=torch.Generator().manual_seed(1)).to(DEV)
    y = mod(x)
    y.backward(gy)
    xo = x.detach().clone().requires_grad_(True)
    layers = [(l.spline_weights.detach().clone().requires_grad_(True), l.linear.weight.detach().clone().requires_grad_(True),
               l.linear.bias.detach().clone().requires_grad_(True)) for l in mod.kan_layers]
    yo = okan.severity_forward(xo, layers, mod.kan_layers[0].knots)
    yo.backward(gy)

    def outlier_rows(a, b, rtol, atol, stol):
        tol = rtol * b.abs() + atol + stol * b.abs().max()
        return ((a - b).abs() > tol).flatten(1).any(dim=1)
    bad_y = outlier_rows(y.detach(), yo.detach(), 1e-3, 1e-5, 0.0)
    bad_dx = outlier_rows(x.grad, xo.grad, 1e-3, 1e-6, 2e-4)
    print(f'\n  y rel-L2 {rel_l2(y, yo):.2e}, dx rel-L2 {rel_l2(x.grad, xo.grad):.2e}; samples out of tolerance: y {int(bad_y.sum())}, dx {int(bad_dx.sum())}')
    assert int((bad_y | bad_dx).sum()) <= OUTLIERS
    assert float((y.detach() - yo.detach()).abs().median()) < 1e-5 and rel_l2(x.grad, xo.grad) < 5e-3
    for li, (l, (sw, lw, lb)) in enumerate(zip(mod.kan_layers, layers)):
        for ours, ref, what in ((l.spline_weights.grad, sw.grad, 'dW'), (l.linear.weight.grad, lw.grad, 'dWl'), (l.linear.bias.grad, lb.grad, 'db')):
            e = rel_l2(ours, ref)
            tol = 1e-3 * ref.abs() + 1e-5 + 5e-4 * ref.abs().max()
            n_bad = int(((ours - ref).abs() > tol).sum())
            print(f'  {l.in_features}->{l.out_features} {what} rel-L2 {e:.2e}, entries out of tolerance {n_bad}/{ref.numel()}')
            assert e < 5e-3
            # a hidden unit whose ReLU gate opens on one side only (the outlier samples above) changes its upstream gradient
            # by O(1) and with it one output column of the first layer's weight gradient: 192 inputs x <= 5 live slots each
            assert n_bad <= OUTLIERS * (192 * 5 if li == 0 else 8), f'{l.in_features}->{l.out_features} {what}'      # measured: 1452 / 11


def test_trunk_against_torchvision_on_device():
    """the trunk oracle is a restatement of timm (timm itself is absent), so the CUDA trunk is ALSO compared
    directly with an independent implementation -- torchvision's VisionTransformer in fp32 on this device -- forward and all
    150 parameter gradients, no oracle code in between."""
    tv = pytest.importorskip('torchvision.models.vision_transformer')
    batch = 33
    sd = _weights(6)
    trunk_sd = {k[len('backbone.model.'):]: v for k, v in sd.items() if k.startswith('backbone.model.')}
    ref = tv.VisionTransformer(image_size=224, patch_size=16, num_layers=12, num_heads=3, hidden_dim=192, mlp_dim=768)
    ref.heads = torch.nn.Identity()
    tv_sd = ovit.to_torchvision(trunk_sd)
    ref.load_state_dict(tv_sd, strict=False)
    ref = ref.to(DEV).eval()
    names = ovit.to_torchvision({k: k for k in trunk_sd})            # torchvision name -> timm name
    images = torch.randn(batch, 3, 224, 224, generator=torch.Generator().manual_seed(6)).to(DEV)
    up = torch.randn(batch, 192, generator=torch.Generator().manual_seed(7)).to(DEV)
    m = _model(sd, train=True)
    f = m.backbone(images)
    (f * up).sum().backward()
    fr = ref(images)
    (fr * up).sum().backward()
    e = rel_l2(f, fr)
    print(f'\n  features rel-L2 vs torchvision {e:.2e}')
    assert e <= OUT_REL_L2
    assert_close(f, fr, rtol=OUT_RTOL, atol=0, scale_tol=OUT_STOL, what='features vs torchvision')
    ours = dict(m.backbone.model.named_parameters())
    errs = {names[k]: rel_l2(ours[names[k]].grad, p.grad.reshape(ours[names[k]].shape)) for k, p in ref.named_parameters() if k in names}
    worst = max((v, k) for k, v in errs.items())
    print('  gradients vs torchvision: worst', worst, 'median', sorted(errs.values())[len(errs) // 2])
    assert len(errs) == 150 and worst[0] <= 2e-2, worst          # measured 7.5e-3 at batch 33
