"""Drop-in for Schrodinger_Equations/Quantum_Harmonic_Oscillator/QHO_2D.py (2-D oscillator, omega = sqrt 2)."""
import math

import torch
import torch.nn as nn

from .. import _lib
from ..ops import ProgramSpec, WanSpec, residual_mean, residual_means, wan_means, wan_scalar_losses
from ._common import Sin, mlp, window_envelope

OMEGA = math.sqrt(2)


def Exact_energy(nx, ny, L):
    """(nx + ny + 1) omega   (QHO_2D.py:93-96)."""
    return (nx + ny + 1) * OMEGA


def hermite_nodes(n):
    """Roots of the n-th oscillator eigenfunction, float32 like the reference (QHO_2D.py:116-143)."""
    s = 2 ** (-1 / 4)
    table = {
        0: [], 1: [0.0], 2: [-2 ** (-3 / 4), 2 ** (-3 / 4)],
        3: [0.0, -2 ** (-3 / 4) * math.sqrt(3), 2 ** (-3 / 4) * math.sqrt(3)],
        4: [-s * math.sqrt((3 + math.sqrt(6)) / 2), -s * math.sqrt((3 - math.sqrt(6)) / 2),
            s * math.sqrt((3 - math.sqrt(6)) / 2), s * math.sqrt((3 + math.sqrt(6)) / 2)],
        5: [0.0, -s * math.sqrt((5 + math.sqrt(10)) / 2), -s * math.sqrt((5 - math.sqrt(10)) / 2),
            s * math.sqrt((5 - math.sqrt(10)) / 2), s * math.sqrt((5 + math.sqrt(10)) / 2)],
    }
    if n not in table:
        raise ValueError(f"Nodes not defined for n={n}")
    return torch.tensor(table[n], dtype=torch.float32)


class FCN(nn.Module):
    """sin network on (x, y) with exp-window envelope and optional nodal lines (QHO_2D.py:103-170)."""

    def __init__(self, layers, nx, ny, technique):
        super().__init__()
        self.nx, self.ny, self.technique = nx, ny, technique
        self.net = mlp(layers, Sin)
        self.nodes_x, self.nodes_y = hermite_nodes(nx), hermite_nodes(ny)

    def forward(self, x, y, L=6.0):
        u = self.net(torch.stack((x, y), dim=-1).view(-1, 2)).view(*x.shape)
        bc = (1 - torch.exp(-(x + L))) * (1 - torch.exp(x - L)) * (1 - torch.exp(-(y + L))) * (1 - torch.exp(y - L))
        if self.technique in ('FBC', 'OG'):
            return bc * u
        if self.technique == 'FN':
            nf = torch.ones_like(x)
            for node in self.nodes_x:
                nf = nf * (x - node.to(x.device))
            for node in self.nodes_y:
                nf = nf * (y - node.to(y.device))
            return bc * nf * u
        raise ValueError(f"Unknown technique: {self.technique}")


def _envelope(model, L):
    t = getattr(model, "technique", None)
    if t in ('FBC', 'OG'):
        return window_envelope(L)
    if t == 'FN':
        return window_envelope(L, [[float(v) for v in model.nodes_x], [float(v) for v in model.nodes_y]])
    raise ValueError(f"Unknown technique: {t}")


def _points(x, y):
    X = torch.stack((x.detach().reshape(-1), y.detach().reshape(-1)), dim=1).contiguous()
    V = 0.5 * OMEGA ** 2 * (X[:, 0] ** 2 + X[:, 1] ** 2)
    return X, V


def PINN_loss(model, x, y, E, L=6.0):
    """mean((-1/2 Lap u + V u - E u)^2): the inline residual block of train_pinn_seperate
    (QHO_2D.py:363-378).  ``E`` may be a float or a trainable scalar tensor (QHO_2D_Energy.py:382-383)."""
    X, V = _points(x, y)
    return residual_mean(model, X, ProgramSpec(_lib.PROG_PINN, alpha=-0.5), _envelope(model, L), beta=V, energy=E)


def DRM_loss(model, x, y, L=6.0):
    """mean(1/2 |grad u|^2 + V u^2) / mean(u^2 + 1e-8)   (QHO_2D.py:381-383)."""
    X, V = _points(x, y)
    m = residual_means(model, X, ProgramSpec(_lib.PROG_RAYLEIGH, alpha=0.5), _envelope(model, L), beta=V)
    return m[0] / (m[1] + 1e-8)


def WAN_loss(u_model, v_model, x, y, nx, ny, L, weight_pde=1.0, weight_norm=1.0, *, u_jets=None, v_jets=None):
    """(total_loss, loss_v, loss_pde, loss_norm)   (QHO_2D.py:204-225)."""
    X, V = _points(x, y)
    m = wan_means(u_model, v_model, X, WanSpec(alpha=0.5, energy_const=Exact_energy(nx, ny, L), w_lo=-float(L), w_hi=float(L),
                                               eps_den=1e-10),
                  env_u=_envelope(u_model, L), env_v=_envelope(v_model, L), beta=V, u_jets=u_jets, v_jets=v_jets)
    loss_pde, loss_v, loss_norm, total_loss = wan_scalar_losses(m, kind=0, vol=4 * L * L, w_pde=weight_pde, w_norm=weight_norm)
    return total_loss, loss_v, loss_pde, loss_norm
