"""Fused optimizer tail (csrc/optimizer.cu, training/optim.py) against torch: GradScaler.unscale_ +
clip_grad_norm_ + torch.optim.AdamW with the reference's two learning-rate groups (training/optimizer.py:18-25), and the
device-side step accounting (N3)."""

import pytest
import torch

from conftest import assert_close

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from rovitkan_b200.models import RoViTKAN
    from rovitkan_b200.training import FusedAdamW, JointLoss, StepStats

DEV = 'cuda'
SHAPES = [(1, 1, 192), (192, 3, 16, 16), (576, 192), (576,), (4097,), (3, 7), (64, 64, 7), (1,)]


def _make(seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(*s, generator=g).to(DEV).requires_grad_(True) for s in SHAPES]


def _grads(step, scale=1.0):
    g = torch.Generator().manual_seed(100 + step)
    return [(torch.randn(*s, generator=g) * (0.3 + 0.2 * i) * scale).to(DEV) for i, s in enumerate(SHAPES)]


def _groups(ps):
    return [{'params': ps[:4], 'lr': 1e-4}, {'params': ps[4:], 'lr': 1e-3}]


@pytest.mark.parametrize('max_norm', [None, 1.0, 1e4])
def test_matches_torch_adamw_with_clipping(max_norm):
    ours, ref = _make(0), _make(0)
    o = FusedAdamW(_groups(ours), weight_decay=1e-4, max_grad_norm=max_norm)
    r = torch.optim.AdamW(_groups(ref), weight_decay=1e-4, foreach=False, fused=False)
    exact = True
    for step in range(3):
        for p, q, g in zip(ours, ref, _grads(step)):
            p.grad, q.grad = g.clone(), g.clone()
        if max_norm is not None:
            want_norm = torch.nn.utils.clip_grad_norm_(ref, max_norm)
        r.step()
        o.step()
        if max_norm is not None:
            assert_close(o.last_grad_norm, want_norm, rtol=1e-5, atol=0, what='global gradient norm')
        for i, (p, q) in enumerate(zip(ours, ref)):
            assert_close(p, q, rtol=2e-6, atol=1e-7, what=f'step {step} tensor {i}')
            exact = exact and torch.equal(p, q)
            assert_close(o.state[p]['exp_avg'], r.state[q]['exp_avg'], rtol=2e-6, atol=1e-9, what='exp_avg')
            assert_close(o.state[p]['exp_avg_sq'], r.state[q]['exp_avg_sq'], rtol=2e-6, atol=1e-12, what='exp_avg_sq')
    print(f'\n  max_norm={max_norm}: bitwise equal to torch.optim.AdamW after 3 steps: {exact}')
    assert float(o.step_count) == 3.0 and float(o.last_step_ran) == 1.0


def test_non_finite_gradient_skips_the_step_like_gradscaler():
    ours = _make(1)
    o = FusedAdamW(_groups(ours), weight_decay=1e-4, max_grad_norm=1.0)
    before = [p.detach().clone() for p in ours]
    gs = _grads(0)
    gs[5][1, 2] = float('inf')
    for p, g in zip(ours, gs):
        p.grad = g
    o.step()
    assert all(torch.equal(a, b) for a, b in zip(before, ours)) and float(o.step_count) == 0.0 and float(o.last_step_ran) == 0.0
    for p, g in zip(ours, _grads(1)):
        p.grad = g
    o.step()
    assert float(o.step_count) == 1.0 and not torch.equal(before[0], ours[0])
    assert torch.isfinite(ours[5]).all()


@pytest.mark.parametrize('flow', ['reference', 'fused'])
def test_gradscaler_flows(flow):
    """'reference' = trainer.py:118-129 (unscale_, clip_grad_norm_, scaler.step); 'fused' = scaler.step alone, with unscale,
    inf check and clipping inside the optimizer kernels.  Both must equal torch AdamW driven the reference way."""
    ours, ref = _make(2), _make(2)
    o = FusedAdamW(_groups(ours), weight_decay=1e-4, max_grad_norm=1.0 if flow == 'fused' else None)
    r = torch.optim.AdamW(_groups(ref), weight_decay=1e-4, foreach=False, fused=False)
    so, sr = torch.amp.GradScaler('cuda', init_scale=1024.0), torch.amp.GradScaler('cuda', init_scale=1024.0)
    sr._lazy_init_scale_growth_tracker(torch.device(DEV))
    so._lazy_init_scale_growth_tracker(torch.device(DEV))
    for step in range(3):
        for p, q, g in zip(ours, ref, _grads(step, scale=1024.0)):
            p.grad, q.grad = g.clone(), g.clone()
        if step == 1:
            ours[2].grad[0, 0] = float('nan')
            ref[2].grad[0, 0] = float('nan')
        # torch side, as the reference's trainer drives it (scale() is emulated by the pre-scaled gradients)
        sr.unscale_(r)
        torch.nn.utils.clip_grad_norm_(ref, 1.0)
        sr.step(r)
        sr.update()
        if flow == 'reference':
            so.unscale_(o)
            torch.nn.utils.clip_grad_norm_(ours, 1.0)
        so.step(o)
        so.update()
        for i, (p, q) in enumerate(zip(ours, ref)):
            assert_close(p, q, rtol=3e-6, atol=1e-7, what=f'{flow} step {step} tensor {i}')
    assert float(so.get_scale()) == float(sr.get_scale()) == 512.0         # the NaN step halved both scales
    assert float(o.step_count) == 2.0


def test_unfreezing_keeps_moments_and_state_dict_round_trip():
    ps = _make(3)
    for p in ps[:4]:
        p.requires_grad_(False)                                  # frozen backbone group (trainer.py:244-246)
    o = FusedAdamW(_groups(ps), weight_decay=1e-4)
    for p, g in zip(ps, _grads(0)):
        p.grad = g if p.requires_grad else None
    o.step()
    m_before = o.state[ps[6]]['exp_avg'].clone()
    for p in ps[:4]:
        p.requires_grad_(True)                                   # epoch 6: unfreeze (trainer.py:62-63)
    for p, g in zip(ps, _grads(1)):
        p.grad = g
    frozen_before = ps[0].detach().clone()
    o.step()
    assert not torch.equal(frozen_before, ps[0]) and float(o.step_count) == 2.0
    assert_close(o.state[ps[6]]['exp_avg'], 0.9 * m_before + 0.1 * _grads(1)[6], rtol=1e-5, atol=2e-7, what='moments kept')
    sd = o.state_dict()
    ps2 = [p.detach().clone().requires_grad_(True) for p in ps]
    o2 = FusedAdamW(_groups(ps2), weight_decay=1e-4)
    o2.load_state_dict(sd)
    for p, q, g in zip(ps, ps2, _grads(2)):
        p.grad, q.grad = g.clone(), g.clone()
    o.step()
    o2.step()
    assert all(torch.equal(p, q) for p, q in zip(ps, ps2))


def test_train_step_with_the_fused_tail_refreshes_the_trunk_weights():
    """The kernels update parameters through raw pointers: the version bump must reach the trunk's bf16 weight shadows."""
    torch.manual_seed(0)
    m = RoViTKAN(pretrained=False, dropout=0.0).to(DEV).train()
    backbone = [p for n, p in m.named_parameters() if 'backbone' in n]
    heads = [p for n, p in m.named_parameters() if 'backbone' not in n]
    opt = FusedAdamW([{'params': backbone, 'lr': 1e-3}, {'params': heads, 'lr': 1e-2}], weight_decay=1e-4, max_grad_norm=1.0)
    x = torch.randn(4, 3, 224, 224, device=DEV)
    y = torch.tensor([0, 1, 2, 3], device=DEV)
    stats = StepStats(DEV)
    feats = []
    for _ in range(3):
        out = m(x)
        losses = JointLoss()(out, y, y, 4)
        opt.zero_grad(set_to_none=True)
        losses['total_loss'].backward()
        opt.step()
        stats.update(losses, out['cls_logits'], y)
        feats.append(out['features'].detach().clone())
    assert not torch.equal(feats[0], feats[1]) and not torch.equal(feats[1], feats[2])
    res = stats.result()
    assert set(res) == {'loss', 'cls_loss', 'ord_loss', 'unc_loss', 'kan_loss', 'accuracy'} and 0.0 <= res['accuracy'] <= 100.0
    assert float(opt.step_count) == 3.0 and torch.isfinite(opt.last_grad_norm)
    m.eval()
    with torch.no_grad():
        f_eval = m(x)['features']
    assert rel(f_eval, feats[2]) > 1e-4                       # eval shadows were rebuilt from the updated weights too


def rel(a, b):
    return float((a - b).norm() / b.norm())


class _NS:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class _TinyConfig:                                              # picklable (the checkpoint stores the config, trainer.py:318)
    def __init__(self, tmp_path, epochs=2, freeze=1, cutmix=True):
        self.flags = _NS(mixed_precision=True, use_cutmix=cutmix, use_mixup=cutmix, cutmix_alpha=1.0, mixup_alpha=0.2,
                         freeze_backbone_epochs=freeze, gradient_clip=1.0, curriculum=True)
        self.train = _NS(epochs=epochs, early_stop_patience=10)
        self.paths = _NS(checkpoints_dir=str(tmp_path))

    def get_stage_for_epoch(self, epoch):
        return min(4, 2 + epoch)                                # stage 3, then 4


def _tiny_config(tmp_path, epochs=2, freeze=1, cutmix=True):
    return _TinyConfig(tmp_path, epochs, freeze, cutmix)


def test_fast_trainer_matches_per_step_item_accounting_and_runs_the_schedule(tmp_path, monkeypatch):
    """FastTrainer (device-side StepStats, fused optimizer tail) against the reference trainer's accounting
    (trainer.py:142-153: .item() every step), then a 2-epoch fit with freeze -> unfreeze, stage switch, CutMix, checkpoint."""
    monkeypatch.setenv('ROVITKAN_SYNTH_PER_CLASS', '6')
    monkeypatch.setenv('ROVITKAN_DATA_WORKERS', '0')
    from rovitkan_b200.data.dataset import DEFAULT_CLASSES, create_dataloaders
    from rovitkan_b200.data.transforms import original_transforms
    from rovitkan_b200.training.fast_trainer import FastTrainer
    sev = {c: i for i, c in enumerate(DEFAULT_CLASSES)}
    tr, va, _ = create_dataloaders(tmp_path / 'a', tmp_path / 'o', DEFAULT_CLASSES, sev, original_transforms(), original_transforms(),
                                   batch_size=8, train_val_split=0.67, num_workers=0, seed=3)
    torch.manual_seed(0)
    m = RoViTKAN(pretrained=False, dropout=0.0).to(DEV)
    groups = [{'params': [p for n, p in m.named_parameters() if 'backbone' in n], 'lr': 0.0},
              {'params': [p for n, p in m.named_parameters() if 'backbone' not in n], 'lr': 0.0}]
    opt = FusedAdamW(groups, weight_decay=0.0, max_grad_norm=1.0)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=2, eta_min=0.0)
    loss_fn = JointLoss(focal_alpha=torch.ones(4, device=DEV))
    cfg = _tiny_config(tmp_path, cutmix=False)
    t = FastTrainer(m, va, va, opt, sched, loss_fn, cfg, torch.device(DEV))          # unshuffled loader, lr = 0: deterministic
    got = t.train_epoch(2)
    # the reference's accounting on the same (unchanged) weights
    m.train()
    m.curriculum_stage = 4
    tot = {k: 0.0 for k in ('total_loss', 'cls_loss', 'ord_loss', 'unc_loss', 'kan_loss')}
    correct = n = 0
    for x, y, s in va:
        x, y, s = x.to(DEV), y.to(DEV), s.to(DEV)
        with torch.autocast('cuda'):
            out = m(x)
            l = loss_fn(out, y, s, 4)
        for k in tot:
            tot[k] += l[k].item()
        correct += out['cls_logits'].max(1)[1].eq(y).sum().item()
        n += y.size(0)
    nb = len(va)
    want = {'loss': tot['total_loss'] / nb, 'cls_loss': tot['cls_loss'] / nb, 'ord_loss': tot['ord_loss'] / nb,
            'unc_loss': tot['unc_loss'] / nb, 'kan_loss': tot['kan_loss'] / nb, 'accuracy': 100.0 * correct / n}
    assert set(got) == set(want)
    for k in want:
        assert abs(got[k] - want[k]) <= 1e-4 * abs(want[k]) + 1e-5, (k, got[k], want[k])
    # ---- the whole schedule with a learning optimizer
    torch.manual_seed(1)
    m2 = RoViTKAN(pretrained=False).to(DEV)
    groups = [{'params': [p for n, p in m2.named_parameters() if 'backbone' in n], 'lr': 1e-5},
              {'params': [p for n, p in m2.named_parameters() if 'backbone' not in n], 'lr': 1e-4}]
    opt2 = FusedAdamW(groups, weight_decay=1e-4, max_grad_norm=1.0)
    sched2 = torch.optim.lr_scheduler.CosineAnnealingLR(opt2, T_max=2, eta_min=1e-6)
    t2 = FastTrainer(m2, tr, va, opt2, sched2, JointLoss(focal_alpha=torch.ones(4, device=DEV)), _tiny_config(tmp_path), torch.device(DEV))
    hist = t2.fit()
    assert len(hist['train_loss']) == 2 and all(v == v and abs(v) < 1e4 for v in hist['train_loss'] + hist['val_loss'])
    assert all(p.requires_grad for p in m2.backbone.parameters())                   # unfrozen at epoch 2
    ck = t2.load_checkpoint(tmp_path / 'best_model.pth')
    assert {'epoch', 'model_state_dict', 'optimizer_state_dict', 'scheduler_state_dict', 'best_val_loss', 'metrics', 'config',
            'scaler_state_dict'} <= set(ck)
    assert float(opt2.step_count) > 0
