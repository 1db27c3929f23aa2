"""Fused optimizer tail: GradScaler unscale + inf check, global-norm gradient clipping and AdamW with the
reference's two learning-rate groups (training/optimizer.py:18-25, training/trainer.py:118-129) in three kernel launches
(csrc/optimizer.cu) instead of ~150 per-tensor kernels; plus device-side step accounting.

`FusedAdamW` is a torch.optim.Optimizer: param_groups, lr schedulers, `zero_grad`, `state_dict` / `load_state_dict` work as
usual.  It can be used exactly like the optimizer the reference builds --

    scaler.unscale_(opt); clip_grad_norm_(...); scaler.step(opt)        # trainer.py:122-128: clipping done outside

-- or with `max_grad_norm=` set, in which case `step()` (or `scaler.step(opt)` WITHOUT a prior unscale_/clip) performs
unscale, inf check, clipping and the update in one pass with no host synchronisation.
"""

from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from .. import _lib


def _bump_versions(tensors):
    """Parameters are updated through raw pointers: tell autograd (and the trunk's weight-shadow cache, ops.EncoderState)."""
    inc = getattr(torch.autograd.graph, 'increment_version', None)
    if inc is not None:
        try:
            inc(tensors)
            return
        except TypeError:
            for t in tensors:
                inc(t)
            return
    torch._foreach_add_(tensors, 0.0)


class FusedAdamW(torch.optim.Optimizer):
    _step_supports_amp_scaling = True       # GradScaler.step hands us grad_scale / found_inf instead of unscaling itself

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 max_grad_norm: Optional[float] = None, grad_mult: float = 1.0):
        if not 0.0 <= lr or not 0.0 <= eps or not 0.0 <= weight_decay:
            raise ValueError('lr, eps and weight_decay must be non-negative')
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        if len(self.param_groups) > 4:
            raise ValueError('FusedAdamW supports up to 4 parameter groups (the reference uses 2)')
        g0 = self.param_groups[0]
        for g in self.param_groups[1:]:
            if (tuple(g['betas']), g['eps'], g['weight_decay']) != (tuple(g0['betas']), g0['eps'], g0['weight_decay']):
                raise ValueError('FusedAdamW: groups may differ in lr only (as in training/optimizer.py:22-25)')
        self.max_grad_norm = max_grad_norm
        self.grad_mult = float(grad_mult)           # e.g. 1 / world_size after a SUM all-reduce
        self._flat = None

    # ---- flat state ---------------------------------------------------------------------------------------------
    def _live_params(self):
        return [(gi, p) for gi, g in enumerate(self.param_groups) for p in g['params'] if p.requires_grad]

    def _build(self):
        live = self._live_params()
        if not live:
            raise RuntimeError('FusedAdamW: no trainable parameters')
        ps = [p for _, p in live]
        for p in ps:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError('FusedAdamW: parameters must be contiguous fp32 CUDA tensors (no CPU fallback)')
        dev = ps[0].device
        numel = (C.c_int64 * len(ps))(*[p.numel() for p in ps])
        n_state = _lib.load().rvk_optimizer_state_floats(len(ps), numel)
        offs, off = [], 0
        for p in ps:
            offs.append(off)
            off += (p.numel() + 4095) // 4096 * 4096
        assert off == n_state
        self._flat = {'params': ps, 'numel': numel, 'groups': (C.c_int * len(ps))(*[gi for gi, _ in live]), 'offsets': offs,
                      'device': dev, 'exp_avg': torch.zeros(n_state, device=dev), 'exp_avg_sq': torch.zeros(n_state, device=dev),
                      'state4': torch.zeros(4, device=dev),
                      'ptable': (C.c_void_p * len(ps))(*[p.data_ptr() for p in ps]),
                      'pkey': tuple(p.data_ptr() for p in ps)}
        for p, o in zip(ps, offs):        # torch-style per-parameter views (state_dict, inspection)
            self.state[p] = {'exp_avg': self._flat['exp_avg'][o:o + p.numel()].view_as(p),
                             'exp_avg_sq': self._flat['exp_avg_sq'][o:o + p.numel()].view_as(p)}
        return self._flat

    def _current(self):
        f = self._flat
        live = [p for _, p in self._live_params()]
        if f is None or len(live) != len(f['params']) or any(a is not b for a, b in zip(live, f['params'])):
            keep, step = None, None
            if f is not None:
                keep = {id(p): (f['exp_avg'][o:o + p.numel()].clone(), f['exp_avg_sq'][o:o + p.numel()].clone())
                        for p, o in zip(f['params'], f['offsets'])}
                step = f['state4'][1:2].clone()
            f = self._build()
            if keep:                       # e.g. backbone unfrozen mid-run (trainer.py:62-63): keep the moments of known tensors
                for p, o in zip(f['params'], f['offsets']):
                    if id(p) in keep:
                        f['exp_avg'][o:o + p.numel()].copy_(keep[id(p)][0])
                        f['exp_avg_sq'][o:o + p.numel()].copy_(keep[id(p)][1])
                f['state4'][1:2].copy_(step)
        pkey = tuple(p.data_ptr() for p in f['params'])
        if pkey != f['pkey']:
            f['ptable'] = (C.c_void_p * len(pkey))(*pkey)
            f['pkey'] = pkey
        return f

    # ---- the step ------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        f = self._current()
        ps = f['params']
        grads = []
        for p in ps:
            g = p.grad
            if g is not None and (g.dtype != torch.float32 or not g.is_contiguous() or g.is_sparse):
                g = g.to_dense().float().contiguous() if g.is_sparse else g.float().contiguous()
                p.grad = g
            grads.append(g)
        if all(g is None for g in grads):
            return loss
        gtable = (C.c_void_p * len(ps))(*[0 if g is None else g.data_ptr() for g in grads])
        lrs = (C.c_double * len(self.param_groups))(*[float(g['lr']) for g in self.param_groups])
        g0 = self.param_groups[0]
        grad_scale = getattr(self, 'grad_scale', None)
        found_inf = getattr(self, 'found_inf', None)
        gs_ptr = fi_ptr = 0
        if grad_scale is not None:
            grad_scale = grad_scale.to(device=f['device'], dtype=torch.float32)
            gs_ptr = grad_scale.data_ptr()
        if found_inf is not None:
            found_inf = found_inf.to(device=f['device'], dtype=torch.float32)
            fi_ptr = found_inf.data_ptr()
        with torch.cuda.device(f['device']):
            _lib.call('rvk_optimizer_step', len(ps), f['ptable'], gtable, f['numel'], f['groups'], f['exp_avg'].data_ptr(),
                      f['exp_avg_sq'].data_ptr(), f['state4'].data_ptr(), lrs, len(self.param_groups), float(g0['betas'][0]),
                      float(g0['betas'][1]), float(g0['eps']), float(g0['weight_decay']),
                      float(self.max_grad_norm) if self.max_grad_norm else 0.0, self.grad_mult, gs_ptr, fi_ptr,
                      torch.cuda.current_stream().cuda_stream)
        _bump_versions([p for p, g in zip(ps, grads) if g is not None])
        return loss

    # ---- device-side results (no host sync unless .item() is called on them) ---------------------------------------
    @property
    def last_grad_norm(self) -> torch.Tensor:
        """Global gradient norm of the last step (after unscaling), as clip_grad_norm_ would have returned it."""
        return self._current()['state4'][2]

    @property
    def step_count(self) -> torch.Tensor:
        return self._current()['state4'][1]

    @property
    def last_step_ran(self) -> torch.Tensor:
        return self._current()['state4'][3]

    def state_dict(self):
        sd = super().state_dict()
        if self._flat is not None:
            sd['fused_step'] = float(self._flat['state4'][1])
        return sd

    def load_state_dict(self, state_dict):
        step = state_dict.get('fused_step')
        super().load_state_dict({k: v for k, v in state_dict.items() if k != 'fused_step'})
        loaded = {p: dict(s) for p, s in self.state.items()}
        self._flat = None
        f = self._build()
        for p, o in zip(f['params'], f['offsets']):
            if p in loaded and 'exp_avg' in loaded[p]:
                f['exp_avg'][o:o + p.numel()].copy_(loaded[p]['exp_avg'].reshape(-1))
                f['exp_avg_sq'][o:o + p.numel()].copy_(loaded[p]['exp_avg_sq'].reshape(-1))
        if step is not None:
            f['state4'][1] = float(step)


class StepStats:
    """Device-side accumulation of what the reference's trainer reads back with six `.item()` calls EVERY step
    (training/trainer.py:144-153: total / cls / ord / unc / kan loss and the number of correct predictions): `update` only
    enqueues adds, `result()` does the single device->host copy."""

    KEYS = ('total_loss', 'cls_loss', 'ord_loss', 'unc_loss', 'kan_loss')

    def __init__(self, device):
        self.acc = torch.zeros(len(self.KEYS) + 1, device=device, dtype=torch.float32)
        self.batches = 0
        self.samples = 0

    @torch.no_grad()
    def update(self, losses: dict, cls_logits: torch.Tensor, class_labels: torch.Tensor):
        vals = torch.stack([losses[k].detach().float().reshape(()) for k in self.KEYS]
                           + [(cls_logits.argmax(dim=1) == class_labels).sum().float()])
        self.acc += vals
        self.batches += 1
        self.samples += int(class_labels.shape[0])

    def result(self) -> dict:
        host = self.acc.cpu().tolist()           # the only synchronisation
        n = max(self.batches, 1)
        out = {('loss' if k == 'total_loss' else k): host[i] / n for i, k in enumerate(self.KEYS)}
        out['accuracy'] = 100.0 * host[-1] / max(self.samples, 1)
        return out
