"""Multi-task heads -- mirror of reference `models/heads.py` on the sm_100a kernels.

Same classes, constructor signatures and parameter names (`fc1`, `fc2`, `fc_mu`, `fc_logvar`) as the
reference (heads.py:7-112).  Each Linear runs through the C ABI with ReLU, dropout and the
log-variance clamp fused into the GEMM epilogue.  The `relu` / `dropout` submodules are kept so the
module tree matches the reference; their work happens inside the fused op.
"""

from typing import Tuple

import torch
import torch.nn as nn

from .._bootstrap import ops as _ops


def _linear(x, lin: nn.Linear, relu: bool = False, drop_p: float = 0.0, clamp=None):
    ops = _ops()
    return ops.LinearFn.apply(*ops.nograd(x, lin.weight, lin.bias), relu, drop_p, clamp)


class _HeadBase(nn.Module):
    def _hidden(self, x: torch.Tensor) -> torch.Tensor:
        p = self.dropout.p if self.training else 0.0
        return _linear(x, self.fc1, relu=True, drop_p=p)


class ClassificationHead(_HeadBase):
    def __init__(self, embed_dim: int = 384, hidden_dim: int = 128, num_classes: int = 4, dropout: float = 0.3):
        super().__init__()
        self.fc1 = nn.Linear(embed_dim, hidden_dim)
        self.relu = nn.ReLU(inplace=True)
        self.dropout = nn.Dropout(dropout)
        self.fc2 = nn.Linear(hidden_dim, num_classes)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return _linear(self._hidden(x), self.fc2)


class OrdinalHead(_HeadBase):
    def __init__(self, embed_dim: int = 384, hidden_dim: int = 128, num_classes: int = 4, dropout: float = 0.3):
        super().__init__()
        self.num_classes = num_classes
        self.num_thresholds = num_classes - 1
        self.fc1 = nn.Linear(embed_dim, hidden_dim)
        self.relu = nn.ReLU(inplace=True)
        self.dropout = nn.Dropout(dropout)
        self.fc2 = nn.Linear(hidden_dim, self.num_thresholds)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return _linear(self._hidden(x), self.fc2)

    @staticmethod
    def probabilities_from_logits(cum_logits: torch.Tensor) -> torch.Tensor:
        """heads.py:45-67 decode: c = sigmoid(logits); p0 = c0, pk = ck - ck-1, pK-1 = 1 - cK-2."""
        c = torch.sigmoid(cum_logits)
        return torch.cat([c[:, :1], c[:, 1:] - c[:, :-1], 1.0 - c[:, -1:]], dim=1)

    def predict_probabilities(self, x: torch.Tensor) -> torch.Tensor:
        return self.probabilities_from_logits(self.forward(x))

    def predict_severity(self, x: torch.Tensor) -> torch.Tensor:
        probs = self.predict_probabilities(x)
        levels = torch.arange(self.num_classes, dtype=torch.float32, device=probs.device)
        return (probs * levels).sum(dim=1, keepdim=True)


class UncertaintyHead(_HeadBase):
    def __init__(self, embed_dim: int = 384, hidden_dim: int = 128, dropout: float = 0.3):
        super().__init__()
        self.fc1 = nn.Linear(embed_dim, hidden_dim)
        self.relu = nn.ReLU(inplace=True)
        self.dropout = nn.Dropout(dropout)
        self.fc_mu = nn.Linear(hidden_dim, 1)
        self.fc_logvar = nn.Linear(hidden_dim, 1)

    def forward(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        h = self._hidden(x)
        mu = _linear(h, self.fc_mu)
        log_var = _linear(h, self.fc_logvar, clamp=(-10.0, 10.0))      # heads.py:100
        return mu, log_var

    def sample(self, x: torch.Tensor, num_samples: int = 100) -> torch.Tensor:
        mu, log_var = self.forward(x)
        std = torch.exp(0.5 * log_var)
        eps = torch.randn(x.size(0), num_samples, device=x.device)
        return mu + std * eps
