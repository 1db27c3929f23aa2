"""Oracle (CPU checker) for the KAN severity path.  TEST INFRASTRUCTURE ONLY.

Restates reference `models/kan.py`:
  * `BSplineBasis.compute_basis`           models/kan.py:10-44
  * `KANLayer.__init__` / `forward`        models/kan.py:48-95
  * `KANSeverityModule.forward`            models/kan.py:138-149

Facts the restatement must keep:
  * 11 uniform knots on [-1, 1] but only `num_knots - degree - 1 = 7` degree-0
    seeds are created (kan.py:12-13, 23-25) and the right-hand recursion term
    is dropped when `i + 1 >= 7` (kan.py:39-40).  The basis is therefore a
    TRUNCATED cubic B-spline family: identically zero for t >= knots[7].
  * the linear branch sees the raw input, the spline branch sees tanh(x)
    (kan.py:74, 92-93); there is no SiLU.
  * the knot buffer is `torch.linspace(-1, 1, 11)` in fp32, so e.g. knots[5] is
    -1.49e-08, not 0; interval membership is decided on those fp32 values.

Three formulations are given so that they can check one another:
  `basis_literal`      the recursion exactly as the reference walks it
  `basis_closed_form`  the per-interval cubic polynomials the CUDA kernel uses
  `layer_forward`      vectorised contraction (einsum) for large batches
and `layer_forward_loop`, which keeps the reference's O(in*out) Python double
loop so that the CPU baseline in bench.py has the reference's cost structure.
"""

from __future__ import annotations

import torch

NUM_KNOTS_DEFAULT = 5
DEGREE_DEFAULT = 3


def make_knots(num_knots: int = NUM_KNOTS_DEFAULT, degree: int = DEGREE_DEFAULT,
               device=None) -> torch.Tensor:
    """Knot buffer of a KANLayer (models/kan.py:59)."""
    return torch.linspace(-1, 1, num_knots + 2 * degree, device=device)


def num_basis(knots: torch.Tensor, degree: int = DEGREE_DEFAULT) -> int:
    return knots.numel() - degree - 1


def basis_literal(t: torch.Tensor, knots: torch.Tensor,
                  degree: int = DEGREE_DEFAULT) -> torch.Tensor:
    """Cox-de Boor recursion with the reference's truncation (kan.py:10-44).

    t: (B, D) fp32, returns (B, D, nb) fp32.
    """
    nk = knots.numel()
    nb = nk - degree - 1
    t = torch.clamp(t, knots[0], knots[-1])                       # kan.py:16
    cur = [((t >= knots[i]) & (t < knots[i + 1])).to(torch.float32)
           for i in range(nb)]                                    # kan.py:23-25
    for d in range(1, degree + 1):                                # kan.py:28
        nxt = []
        for i in range(nb):
            acc = torch.zeros_like(t, dtype=torch.float32)
            den_l = knots[i + d] - knots[i]
            if float(den_l) != 0.0:                               # kan.py:32
                acc = acc + (t - knots[i]) / den_l * cur[i]
            if i + d + 1 < nk:                                    # kan.py:37
                den_r = knots[i + d + 1] - knots[i + 1]
                if float(den_r) != 0.0 and i + 1 < nb:            # kan.py:37,39
                    acc = acc + (knots[i + d + 1] - t) / den_r * cur[i + 1]
            nxt.append(acc)
        cur = nxt
    return torch.stack(cur, dim=-1)


def basis_closed_form(t: torch.Tensor, knots: torch.Tensor,
                      with_derivative: bool = False):
    """Per-interval cubic form of the truncated basis (degree 3, uniform knots).

    With j the interval index (knots[j] <= t < knots[j+1]) and
    u = (t - knots[j]) / (knots[j+1] - knots[j]) the four live functions are
        N_j   = u^3 / 6
        N_j-1 = (1 + 3u + 3u^2 - 3u^3) / 6
        N_j-2 = (4 - 6u^2 + 3u^3) / 6
        N_j-3 = (1 - u)^3 / 6
    (indices outside [0, nb) dropped), and everything is zero when j >= nb,
    which is where the reference's missing degree-0 seeds bite.  This is the
    formulation the CUDA kernel evaluates; the test-suite checks it against
    `basis_literal` and against the reference's own `compute_basis` vectors.
    """
    nk = knots.numel()
    nb = nk - 3 - 1
    tt = torch.clamp(t, knots[0], knots[-1])
    # interval index on the fp32 knot values themselves
    j = torch.zeros_like(tt, dtype=torch.long) - 1
    for i in range(nk - 1):
        inside = (tt >= knots[i]) & (tt < knots[i + 1])
        j = torch.where(inside, torch.full_like(j, i), j)
    live = (j >= 0) & (j < nb)
    jc = j.clamp(0, nk - 2)
    k0 = knots[jc]
    h = knots[jc + 1] - k0
    u = (tt - k0) / h
    u2, u3 = u * u, u * u * u
    vals = [u3 / 6, (1 + 3 * u + 3 * u2 - 3 * u3) / 6,
            (4 - 6 * u2 + 3 * u3) / 6, (1 - u) ** 3 / 6]
    ders = [u2 / 2, (3 + 6 * u - 9 * u2) / 6, (-12 * u + 9 * u2) / 6,
            -((1 - u) ** 2) / 2]
    out = torch.zeros(*t.shape, nb, dtype=torch.float32, device=t.device)
    dout = torch.zeros_like(out)
    for m in range(4):
        idx = j - m
        ok = live & (idx >= 0) & (idx < nb)
        sel = idx.clamp(0, nb - 1).unsqueeze(-1)
        out.scatter_add_(-1, sel, torch.where(ok, vals[m], torch.zeros_like(u)).unsqueeze(-1))
        dout.scatter_add_(-1, sel, torch.where(ok, ders[m] / h, torch.zeros_like(u)).unsqueeze(-1))
    if with_derivative:
        return out, dout
    return out


def layer_forward(x: torch.Tensor, spline_weights: torch.Tensor,
                  lin_weight: torch.Tensor, lin_bias: torch.Tensor,
                  knots: torch.Tensor, degree: int = DEGREE_DEFAULT) -> torch.Tensor:
    """KANLayer.forward (kan.py:70-95), contraction vectorised.

    x (B, in) -> (B, out).  spline_weights is (in, out, nb) as in the reference.
    Differentiable through torch autograd (used for gradient parity).
    """
    basis = basis_literal(torch.tanh(x), knots, degree)           # kan.py:74-79
    spline = torch.einsum('bik,iok->bo', basis, spline_weights)   # kan.py:83-89
    return torch.nn.functional.linear(x, lin_weight, lin_bias) + spline  # kan.py:92-93


def layer_forward_loop(x: torch.Tensor, spline_weights: torch.Tensor,
                       lin_weight: torch.Tensor, lin_bias: torch.Tensor,
                       knots: torch.Tensor, degree: int = DEGREE_DEFAULT) -> torch.Tensor:
    """Same result as `layer_forward`, accumulated one (input, output) pair at
    a time in the order the reference does (kan.py:85-89).  O(in*out) small
    ops: this is what makes the reference slow, and it is what the CPU
    baseline times."""
    n_in, n_out, _ = spline_weights.shape
    basis = basis_literal(torch.tanh(x), knots, degree)
    spline = torch.zeros(x.shape[0], n_out, dtype=torch.float32, device=x.device)
    for i in range(n_in):
        for o in range(n_out):
            # same op sequence per pair as the reference: slice, multiply, reduce, in-place column add
            spline[:, o] += (basis[:, i, :] * spline_weights[i, o]).sum(dim=1)
    return torch.nn.functional.linear(x, lin_weight, lin_bias) + spline


def severity_forward(x: torch.Tensor, layers, knots: torch.Tensor,
                     degree: int = DEGREE_DEFAULT, loop: bool = False,
                     return_trajectory: bool = False):
    """KANSeverityModule.forward (kan.py:138-149).

    `layers` is a list of (spline_weights, lin_weight, lin_bias).  ReLU between
    layers, 3*sigmoid on the last one.
    """
    fn = layer_forward_loop if loop else layer_forward
    traj = [x]
    for li, (sw, lw, lb) in enumerate(layers):
        x = fn(x, sw, lw, lb, knots, degree)
        if li < len(layers) - 1:
            x = torch.relu(x)
        else:
            x = 3.0 * torch.sigmoid(x)
        traj.append(x)
    return traj if return_trajectory else x


def init_layers(dims, num_knots: int = NUM_KNOTS_DEFAULT, degree: int = DEGREE_DEFAULT,
                generator: torch.Generator | None = None):
    """Random parameters with the reference's shapes and init laws
    (kan.py:63-68: spline ~ 0.1*N(0,1); nn.Linear default init)."""
    nb = num_knots + degree - 1
    out = []
    for a, b in zip(dims[:-1], dims[1:]):
        sw = torch.randn(a, b, nb, generator=generator) * 0.1
        bound = 1.0 / (a ** 0.5)
        lw = (torch.rand(b, a, generator=generator) * 2 - 1) * bound
        lb = (torch.rand(b, generator=generator) * 2 - 1) * bound
        out.append((sw, lw, lb))
    return out
