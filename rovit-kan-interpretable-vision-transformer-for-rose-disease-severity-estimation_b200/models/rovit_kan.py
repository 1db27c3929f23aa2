"""RoViT-KAN -- mirror of reference `models/rovit_kan.py` (rovit_kan.py:9-181) on the sm_100a kernels.

Same constructor (first argument a Config, an int or None; additionally accepts `embed_dim=` because
the reference's own scripts call it that way, scripts/train.py:88-97), same attributes
(`backbone`, `classification_head`, `ordinal_head`, `uncertainty_head`, `kan_module`,
`curriculum_stage`), same output dict keys and stage gating, same state_dict keys.
"""

from typing import Dict

import torch
import torch.nn as nn

from .._bootstrap import ops as _ops
from .backbone import DeiTTinyBackbone
from .heads import ClassificationHead, OrdinalHead, UncertaintyHead
from .kan import KANSeverityModule


def _serial_heads() -> bool:
    import os
    return os.environ.get('RVK_SERIAL_HEADS', '0') == '1'


def _fused_train_tail() -> bool:
    import os
    return os.environ.get('RVK_FUSED_TAIL', '1') != '0'


class RoViTKAN(nn.Module):
    def __init__(self, config_or_embed_dim=None, hidden_dim: int = 128, num_classes: int = 4,
                 kan_layers: list = None, kan_num_knots: int = 5, kan_degree: int = 3, dropout: float = 0.3,
                 pretrained: bool = True, embed_dim: int = None):
        super().__init__()
        if hasattr(config_or_embed_dim, 'model'):
            cfg = config_or_embed_dim
            embed_dim = cfg.model.embed_dim
            hidden_dim = cfg.model.hidden_dim
            num_classes = cfg.data.num_classes
            kan_layers = cfg.model.kan_layers
            kan_num_knots = cfg.model.kan_num_knots
            kan_degree = cfg.model.kan_degree
            dropout = cfg.model.dropout
            pretrained = cfg.model.pretrained
        elif config_or_embed_dim is not None:
            embed_dim = config_or_embed_dim

        self.backbone = DeiTTinyBackbone(pretrained=pretrained, freeze=False)
        if embed_dim is None:
            embed_dim = self.backbone.embed_dim
        if kan_layers is None:
            kan_layers = [embed_dim, 64, 16, 1]

        self.classification_head = ClassificationHead(embed_dim=embed_dim, hidden_dim=hidden_dim,
                                                      num_classes=num_classes, dropout=dropout)
        self.ordinal_head = OrdinalHead(embed_dim=embed_dim, hidden_dim=hidden_dim, num_classes=num_classes,
                                        dropout=dropout)
        self.uncertainty_head = UncertaintyHead(embed_dim=embed_dim, hidden_dim=hidden_dim, dropout=dropout)
        self.kan_module = KANSeverityModule(layers=kan_layers, num_knots=kan_num_knots, degree=kan_degree)
        self._curriculum_stage = 4

    @property
    def curriculum_stage(self) -> int:
        return self._curriculum_stage

    @curriculum_stage.setter
    def curriculum_stage(self, stage: int):
        assert 1 <= stage <= 4, "Stage must be between 1 and 4"
        self._curriculum_stage = stage

    def _fused_tail_params(self):
        """The 23 head / KAN tensors in the order rvk_heads_fused expects, or None if this instance does not have the
        reference architecture (192 -> 128 -> {4,3,1,1}, kan_layers [192,64,16,1])."""
        c, o, u, k = self.classification_head, self.ordinal_head, self.uncertainty_head, self.kan_module
        if (tuple(c.fc1.weight.shape) != (128, 192) or tuple(c.fc2.weight.shape) != (4, 128) or
                tuple(o.fc2.weight.shape) != (3, 128) or list(k.layers_dims) != [192, 64, 16, 1]):
            return None
        if len({l.knots_host() for l in k.kan_layers}) != 1:      # the fused kernel uses one knot vector for the stack
            return None
        ps = [c.fc1.weight, c.fc1.bias, c.fc2.weight, c.fc2.bias, o.fc1.weight, o.fc1.bias, o.fc2.weight, o.fc2.bias,
              u.fc1.weight, u.fc1.bias, u.fc_mu.weight, u.fc_mu.bias, u.fc_logvar.weight, u.fc_logvar.bias]
        for l in k.kan_layers:
            ps += [l.spline_weights, l.linear.weight, l.linear.bias]
        return ps

    def forward(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        features = self.backbone(x)
        stage = self._curriculum_stage
        if stage == 4 and not self.training and not torch.is_grad_enabled() and features.is_cuda:
            # inference: the whole multi-task tail is one kernel launch (csrc/heads_fused.cuh)
            ps = self._fused_tail_params()
            if ps is not None:
                ops = _ops()
                if getattr(self, '_tail_state', None) is None:
                    self._tail_state = ops.HeadsFusedState()
                cls, ordl, mu, lv, kan = ops.heads_fused(self._tail_state, features, ps, self.kan_module.kan_layers[0].knots_host())
                return {'cls_logits': cls, 'features': features, 'ordinal_logits': ordl, 'mu': mu, 'log_var': lv,
                        'kan_severity': kan}
        out = {'cls_logits': None, 'features': features, 'ordinal_logits': None, 'mu': None, 'log_var': None, 'kan_severity': None}
        if torch.is_grad_enabled() and features.is_cuda and _fused_train_tail():
            # training: the whole tail is one forward and one backward kernel (csrc/heads_fused.cuh<true>, heads_train.cuh);
            # gated outputs are simply not handed out, so their parameters receive no gradient
            ps = self._fused_tail_params()
            drops = {h.dropout.p if self.training else 0.0 for h in (self.classification_head, self.ordinal_head, self.uncertainty_head)}
            if ps is not None and len(drops) == 1:
                cls, ordl, mu, lv, kan = _ops().HeadsTrainFn.apply(features, self.kan_module.kan_layers[0].knots_host(), drops.pop(), *ps)
                out['cls_logits'] = cls
                if stage >= 2:
                    out['ordinal_logits'] = ordl
                if stage >= 3:
                    out['mu'], out['log_var'] = mu, lv
                if stage >= 4:
                    out['kan_severity'] = kan
                return out
        branches = [('cls', self.classification_head)]
        if stage >= 2:
            branches.append(('ord', self.ordinal_head))
        if stage >= 3:
            branches.append(('unc', self.uncertainty_head))
        if stage >= 4:
            branches.append(('kan', self.kan_module))
        res = self._run_branches(features, branches)
        out['cls_logits'] = res['cls']
        if 'ord' in res:
            out['ordinal_logits'] = res['ord']
        if 'unc' in res:
            out['mu'], out['log_var'] = res['unc']
        if 'kan' in res:
            out['kan_severity'] = res['kan']
        return out

    def _run_branches(self, features, branches):
        """The heads are independent given the features, and at training batch sizes each of them is a chain of small,
        latency-bound kernels: run every branch on its own CUDA stream (fork after the trunk, join before the loss).  autograd
        replays each branch's backward on the stream its forward ran on, so the backward chains overlap the same way.
        RVK_SERIAL_HEADS=1 keeps everything on the current stream."""
        if len(branches) == 1 or not features.is_cuda or _serial_heads():
            return {name: fn(features) for name, fn in branches}
        cur = torch.cuda.current_stream(features.device)
        streams = getattr(self, '_branch_streams', None)
        if streams is None or len(streams) < len(branches) - 1 or streams[0].device != features.device:
            streams = [torch.cuda.Stream(device=features.device) for _ in range(3)]
            self._branch_streams = streams
            # the head parameters' AccumulateGrad nodes live on the default stream while their gradients are produced on the
            # branch streams: intentional (autograd synchronises them), so silence the advisory warning about it
            quiet = getattr(torch.autograd.graph, 'set_warn_on_accumulate_grad_stream_mismatch', None)
            if quiet is not None:
                quiet(False)
        fork = torch.cuda.Event()
        fork.record(cur)
        res = {}
        name0, fn0 = branches[-1]                  # the longest chain (KAN stack in stage 4) stays on the current stream
        for i, (name, fn) in enumerate(branches[:-1]):
            s = streams[i]
            s.wait_event(fork)
            with torch.cuda.stream(s):
                res[name] = fn(features)
            features.record_stream(s)
        res[name0] = fn0(features)
        for i in range(len(branches) - 1):
            cur.wait_stream(streams[i])
        for name, _ in branches[:-1]:              # produced on a side stream, consumed (loss) on the current one
            for t in (res[name] if isinstance(res[name], tuple) else (res[name],)):
                t.record_stream(cur)
        return res

    def predict(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        self.eval()
        with torch.no_grad():
            out = self.forward(x)
            # one decode launch (csrc/heads.cu::predict_decode_kernel).  The reference re-runs the ordinal head twice here
            # (rovit_kan.py:143-148); in eval mode that recomputes the same logits, so the ones already in hand are decoded
            idx, probs, oprobs, osev, std = _ops().predict_decode(out['cls_logits'], out['ordinal_logits'], out['log_var'])
            pred = {'class': idx, 'class_probs': probs, 'features': out['features']}
            if out['ordinal_logits'] is not None:
                pred['ordinal_probs'] = oprobs
                pred['ordinal_severity'] = osev
            if out['mu'] is not None:
                pred['uncertainty_mu'] = out['mu']
                pred['uncertainty_std'] = std
            if out['kan_severity'] is not None:
                pred['kan_severity'] = out['kan_severity']
            return pred

    def freeze_backbone(self):
        self.backbone.freeze()

    def unfreeze_backbone(self):
        self.backbone.unfreeze()

    def get_attention_maps(self, x: torch.Tensor):
        return self.backbone.get_attention_maps(x)

    def count_parameters(self) -> Dict[str, int]:
        n = lambda m: sum(p.numel() for p in m.parameters() if p.requires_grad)
        counts = {'backbone': n(self.backbone), 'classification_head': n(self.classification_head),
                  'ordinal_head': n(self.ordinal_head), 'uncertainty_head': n(self.uncertainty_head),
                  'kan_module': n(self.kan_module)}
        counts['total'] = sum(counts.values())
        return counts
