"""Joint loss -- mirror of reference `training/losses.py` (losses.py:7-181) on one fused CUDA kernel.

Same classes and constructor signatures.  `JointLoss.forward` evaluates all four terms, their batch
means and the local gradients in a single kernel (csrc/heads.cu::joint_loss_kernel) and returns the
reference's dict (`cls_loss, ord_loss, unc_loss, kan_loss, total_loss`).  The per-term classes route
through the same kernel.  Only reduction='mean' (the only mode the reference's trainer uses) is built.
"""

from typing import Dict, Optional

import torch
import torch.nn as nn

from .._bootstrap import ops as _ops


def _check_reduction(reduction: str):
    if reduction != 'mean':
        raise NotImplementedError("only reduction='mean' is built into the fused loss kernel")


def _joint(cls_logits, ord_logits, mu, log_var, kan, class_t, sev_t, alpha, gamma, lam, mu_unc, nu_kan):
    return _ops().JointLossFn.apply(cls_logits, ord_logits, mu, log_var, kan, class_t, sev_t, alpha, gamma, lam,
                                    mu_unc, nu_kan)


class FocalLoss(nn.Module):
    def __init__(self, gamma: float = 2.0, alpha: Optional[torch.Tensor] = None, reduction: str = 'mean'):
        super().__init__()
        _check_reduction(reduction)
        self.gamma, self.alpha, self.reduction = gamma, alpha, reduction

    def forward(self, logits: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        return _joint(logits, None, None, None, None, targets, targets, self.alpha, self.gamma, 0.0, 0.0, 0.0)[0]


class OrdinalBCELoss(nn.Module):
    def __init__(self, num_classes: int = 4, reduction: str = 'mean'):
        super().__init__()
        _check_reduction(reduction)
        self.num_classes, self.num_thresholds, self.reduction = num_classes, num_classes - 1, reduction

    def forward(self, cum_logits: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        dummy = torch.zeros(cum_logits.shape[0], self.num_classes, device=cum_logits.device)
        zeros = torch.zeros(targets.shape, dtype=torch.int64, device=targets.device)
        return _joint(dummy, cum_logits, None, None, None, zeros, targets, None, 2.0, 1.0, 0.0, 0.0)[1]


class UncertaintyLoss(nn.Module):
    def __init__(self, reduction: str = 'mean'):
        super().__init__()
        _check_reduction(reduction)
        self.reduction = reduction

    def forward(self, mu: torch.Tensor, log_var: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        dummy = torch.zeros(mu.shape[0], 2, device=mu.device)
        t = targets.reshape(-1).float()
        return _joint(dummy, None, mu, log_var, None, torch.zeros_like(t, dtype=torch.int64), t, None, 2.0, 0.0, 1.0, 0.0)[2]


class KANRegressionLoss(nn.Module):
    def __init__(self, reduction: str = 'mean'):
        super().__init__()
        _check_reduction(reduction)
        self.reduction = reduction

    def forward(self, predictions: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        dummy = torch.zeros(predictions.shape[0], 2, device=predictions.device)
        t = targets.reshape(-1).float()
        return _joint(dummy, None, None, None, predictions, torch.zeros_like(t, dtype=torch.int64), t, None, 2.0, 0.0, 0.0, 1.0)[3]


class JointLoss(nn.Module):
    def __init__(self, lambda_ord: float = 1.0, mu_unc: float = 0.5, nu_kan: float = 0.5, focal_gamma: float = 2.0,
                 focal_alpha: Optional[torch.Tensor] = None, num_classes: int = 4):
        super().__init__()
        self.lambda_ord, self.mu_unc, self.nu_kan = lambda_ord, mu_unc, nu_kan
        self.focal_loss = FocalLoss(gamma=focal_gamma, alpha=focal_alpha)
        self.ordinal_loss = OrdinalBCELoss(num_classes=num_classes)
        self.uncertainty_loss = UncertaintyLoss()
        self.kan_loss = KANRegressionLoss()
        self._alpha_dev = None

    def _alpha(self, device):
        a = self.focal_loss.alpha
        if a is None:
            return None
        if self._alpha_dev is None or self._alpha_dev.device != device or self._alpha_dev.shape != a.shape:
            self._alpha_dev = a.detach().to(device=device, dtype=torch.float32).contiguous()
        return self._alpha_dev

    def forward(self, outputs: Dict[str, torch.Tensor], class_targets: torch.Tensor,
                severity_targets: torch.Tensor, stage: int = 4) -> Dict[str, torch.Tensor]:
        cls_logits = outputs['cls_logits']
        ordl = outputs['ordinal_logits'] if stage >= 2 else None
        mu = outputs['mu'] if stage >= 3 else None
        lv = outputs['log_var'] if stage >= 3 else None
        if mu is None or lv is None:
            mu = lv = None
        kan = outputs['kan_severity'] if stage >= 4 else None
        out = _joint(cls_logits, ordl, mu, lv, kan, class_targets, severity_targets, self._alpha(cls_logits.device),
                     self.focal_loss.gamma, self.lambda_ord, self.mu_unc, self.nu_kan)
        return {'cls_loss': out[0], 'ord_loss': out[1], 'unc_loss': out[2], 'kan_loss': out[3], 'total_loss': out[4]}
