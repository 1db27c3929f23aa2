"""Oracle (CPU checker) for the joint loss.  TEST INFRASTRUCTURE ONLY.

Restates reference `training/losses.py`:
  FocalLoss.forward          losses.py:15-38    alpha[y] * (1-pt)^gamma * CE, mean
  OrdinalBCELoss.forward     losses.py:48-72    BCE-with-logits on (y > k), mean over (B,3)
  UncertaintyLoss.forward    losses.py:80-101   0.5*((y-mu)^2*exp(-lv) + lv), mean
  KANRegressionLoss.forward  losses.py:109-114  mse(pred, y.float()[:,None])
  JointLoss.forward          losses.py:139-181  cls + l_ord*ord + mu_unc*unc + nu_kan*kan,
                                                each gated by `stage` and by the output
                                                being present; absent terms are 0.
Note: the ordinal loss trains sigmoid(logit_k) towards [y > k] while
`predict_probabilities` reads it as P(y <= k).  Both are reproduced as written.
"""

from __future__ import annotations

import torch
import torch.nn.functional as F


def focal(logits, targets, gamma: float = 2.0, alpha=None):
    logp = F.log_softmax(logits, dim=1)
    logpt = logp.gather(1, targets[:, None]).squeeze(1)
    ce = -logpt
    pt = torch.softmax(logits, dim=1).gather(1, targets[:, None]).squeeze(1)
    loss = (1 - pt) ** gamma * ce
    if alpha is not None:
        loss = alpha.to(logits.device)[targets] * loss
    return loss.mean()


def ordinal_bce(cum_logits, targets):
    k = torch.arange(cum_logits.shape[1], device=cum_logits.device)
    tgt = (targets[:, None] > k[None, :]).to(torch.float32)
    per = F.binary_cross_entropy_with_logits(cum_logits, tgt, reduction='none')
    return per.mean(dim=1).mean()


def uncertainty_nll(mu, log_var, targets):
    y = targets.to(torch.float32)[:, None]
    return (0.5 * ((y - mu) ** 2 * torch.exp(-log_var) + log_var)).mean()


def kan_mse(pred, targets):
    return F.mse_loss(pred, targets.to(torch.float32)[:, None])


def joint(outputs: dict, class_targets, severity_targets, stage: int = 4,
          lambda_ord: float = 1.0, mu_unc: float = 0.5, nu_kan: float = 0.5,
          gamma: float = 2.0, alpha=None) -> dict:
    zero = lambda: torch.tensor(0.0, device=outputs['cls_logits'].device)
    out = {}
    out['cls_loss'] = focal(outputs['cls_logits'], class_targets, gamma, alpha)
    total = out['cls_loss']
    if stage >= 2 and outputs.get('ordinal_logits') is not None:
        out['ord_loss'] = ordinal_bce(outputs['ordinal_logits'], severity_targets)
        total = total + lambda_ord * out['ord_loss']
    else:
        out['ord_loss'] = zero()
    if stage >= 3 and outputs.get('mu') is not None and outputs.get('log_var') is not None:
        out['unc_loss'] = uncertainty_nll(outputs['mu'], outputs['log_var'], severity_targets)
        total = total + mu_unc * out['unc_loss']
    else:
        out['unc_loss'] = zero()
    if stage >= 4 and outputs.get('kan_severity') is not None:
        out['kan_loss'] = kan_mse(outputs['kan_severity'], severity_targets)
        total = total + nu_kan * out['kan_loss']
    else:
        out['kan_loss'] = zero()
    out['total_loss'] = total
    return out
