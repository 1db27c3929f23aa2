"""KAN severity regressor -- mirror of reference `models/kan.py` on the fused sm_100a kernels.

Same classes, constructor signatures, parameter/buffer names and shapes as the reference
(`BSplineBasis`, `KANLayer`, `KANSeverityModule`; kan.py:8-170); `forward` calls the fused CUDA
KAN layer (csrc/kan.cu) through the C ABI instead of the reference's Python double loop
(kan.py:85-89).  CUDA only.
"""

from typing import List, Tuple

import numpy as np
import torch
import torch.nn as nn

from .._bootstrap import ops as _ops

_DEGREE = 3


def _check_config(num_knots: int, degree: int):
    if degree != _DEGREE or num_knots != 5:
        raise NotImplementedError(
            f'the sm_100a KAN kernels are built for the reference configuration num_knots=5, degree=3 '
            f'(11 knots, 7 basis functions); got num_knots={num_knots}, degree={degree}')


class BSplineBasis:
    @staticmethod
    def compute_basis(x: torch.Tensor, knots: torch.Tensor, degree: int = 3) -> torch.Tensor:
        """Truncated Cox-de Boor basis of the reference (kan.py:10-44): x (B, D) already in knot range
        -> (B, D, num_knots_total - degree - 1)."""
        if degree != _DEGREE or knots.numel() != 11:
            raise NotImplementedError('only degree 3 on 11 knots (reference configuration) is built')
        return _ops().kan_basis(x, [float(v) for v in knots.detach().cpu().tolist()])


class KANLayer(nn.Module):
    def __init__(self, in_features: int, out_features: int, num_knots: int = 5, degree: int = 3):
        super().__init__()
        _check_config(num_knots, degree)
        self.in_features = in_features
        self.out_features = out_features
        self.num_knots = num_knots
        self.degree = degree
        self.num_basis = num_knots + degree - 1
        # same construction order and init laws as the reference (kan.py:58-68)
        self.register_buffer('knots', torch.linspace(-1, 1, num_knots + 2 * degree))
        self.spline_weights = nn.Parameter(torch.randn(in_features, out_features, self.num_basis) * 0.1)
        self.linear = nn.Linear(in_features, out_features, bias=True)
        self._knots_host = None
        self._knots_version = None

    def knots_host(self):
        """Host copy of the knot buffer (refreshed if the buffer is replaced or mutated)."""
        key = (self.knots.data_ptr(), self.knots._version)
        if self._knots_host is None or self._knots_version != key:
            self._knots_host = tuple(float(v) for v in self.knots.detach().cpu().tolist())
            self._knots_version = key
        return self._knots_host

    def _forward_act(self, x: torch.Tensor, act: int) -> torch.Tensor:
        ops = _ops()
        if getattr(self, '_ws_state', None) is None:
            self._ws_state = ops.KanLayerState()
        return ops.KanLayerFn.apply(*ops.nograd(x, self.spline_weights, self.linear.weight, self.linear.bias),
                                    self.knots_host(), act, self._ws_state)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self._forward_act(x, 0)

    def get_spline_weights(self) -> torch.Tensor:
        return self.spline_weights.detach()

    def plot_activation(self, input_idx: int = 0, output_idx: int = 0,
                        num_points: int = 100) -> Tuple[np.ndarray, np.ndarray]:
        x_vals = torch.linspace(-1, 1, num_points, device=self.knots.device)
        basis = BSplineBasis.compute_basis(x_vals[None, :], self.knots, self.degree)      # (1, P, nb)
        y_vals = (basis[0] * self.spline_weights[input_idx, output_idx].detach()).sum(dim=1)
        return x_vals.cpu().numpy(), y_vals.cpu().numpy()


class KANSeverityModule(nn.Module):
    def __init__(self, layers: List[int] = [384, 64, 16, 1], num_knots: int = 5, degree: int = 3):
        super().__init__()
        self.layers_dims = layers
        self.num_knots = num_knots
        self.degree = degree
        self.kan_layers = nn.ModuleList(
            [KANLayer(layers[i], layers[i + 1], num_knots, degree) for i in range(len(layers) - 1)])
        self.activations = nn.ModuleList([nn.ReLU() for _ in range(len(layers) - 2)])

    def _run(self, x: torch.Tensor, keep: bool):
        acts = [x]
        last = len(self.kan_layers) - 1
        for i, layer in enumerate(self.kan_layers):
            # ReLU between layers and 3*sigmoid at the end are fused into the layer kernel's epilogue
            x = layer._forward_act(x, 2 if i == last else 1)
            if keep:
                acts.append(x)
        return x, acts

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self._run(x, False)[0]

    def get_spline_weights(self) -> List[torch.Tensor]:
        return [layer.get_spline_weights() for layer in self.kan_layers]

    def get_activation_trajectory(self, x: torch.Tensor) -> List[torch.Tensor]:
        return self._run(x, True)[1]

    def count_parameters(self) -> int:
        return sum(p.numel() for p in self.parameters() if p.requires_grad)
