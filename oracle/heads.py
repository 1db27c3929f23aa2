"""Oracle (CPU checker) for the three MLP heads.  TEST INFRASTRUCTURE ONLY.

Restates reference `models/heads.py`:
  ClassificationHead.forward        heads.py:17-22   fc2(drop(relu(fc1 x)))
  OrdinalHead.forward               heads.py:38-43   same, 3 cumulative logits
  OrdinalHead.predict_probabilities heads.py:45-67   c=sigmoid; p0=c0, pk=ck-ck-1, p3=1-c2
  OrdinalHead.predict_severity      heads.py:69-77   sum_k k*p_k, keepdim
  UncertaintyHead.forward           heads.py:91-102  mu, clamp(log_var, -10, 10)

Dropout is the identity here (eval mode, or p=0 for gradient parity): the
reference's RNG stream cannot be reproduced bit-for-bit by a fused kernel, so
train-mode dropout is covered by a statistical test instead.
An optional `keep` mask (already scaled by 1/(1-p)) lets tests replay the mask
the CUDA kernel drew.
"""

from __future__ import annotations

import torch
import torch.nn.functional as F


def mlp_hidden(x, fc1_w, fc1_b, keep=None):
    h = torch.relu(F.linear(x, fc1_w, fc1_b))
    if keep is not None:
        h = h * keep
    return h


def classification_forward(x, fc1_w, fc1_b, fc2_w, fc2_b, keep=None):
    return F.linear(mlp_hidden(x, fc1_w, fc1_b, keep), fc2_w, fc2_b)


ordinal_forward = classification_forward   # same shape of computation, 3 outputs


def ordinal_probabilities(cum_logits: torch.Tensor) -> torch.Tensor:
    c = torch.sigmoid(cum_logits)
    first = c[:, :1]
    mid = c[:, 1:] - c[:, :-1]
    last = 1.0 - c[:, -1:]
    return torch.cat([first, mid, last], dim=1)


def ordinal_severity(cum_logits: torch.Tensor) -> torch.Tensor:
    p = ordinal_probabilities(cum_logits)
    levels = torch.arange(p.shape[1], dtype=torch.float32, device=p.device)
    return (p * levels).sum(dim=1, keepdim=True)


def uncertainty_forward(x, fc1_w, fc1_b, mu_w, mu_b, lv_w, lv_b, keep=None):
    h = mlp_hidden(x, fc1_w, fc1_b, keep)
    mu = F.linear(h, mu_w, mu_b)
    log_var = torch.clamp(F.linear(h, lv_w, lv_b), min=-10, max=10)
    return mu, log_var
