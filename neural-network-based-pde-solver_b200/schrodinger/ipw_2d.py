"""Drop-in for Schrodinger_Equations/Infinite_Potential_Well/IPW_2D.py (2-D infinite well)."""
import numpy as np
import torch
import torch.nn as nn

from .. import _lib
from ..ops import ProgramSpec, residual_mean, residual_means
from ._common import Sin, mlp, poly_envelope, values


def Exact_solution(L, nx, ny, x, y):
    """(2 / L) sin(nx pi x / L) sin(ny pi y / L)   (IPW_2D.py:69-71)."""
    return (2.0 / L) * torch.sin(nx * torch.pi * x / L) * torch.sin(ny * torch.pi * y / L)


class FCN(nn.Module):
    """sin network on (x, y) with x (L - x) y (L - y) and optional nodal lines (IPW_2D.py:78-110)."""

    def __init__(self, layers, nx, ny, technique):
        super().__init__()
        self.nx, self.ny, self.technique = nx, ny, technique
        self.net = mlp(layers, Sin)

    def forward(self, x, y, L=2.0):
        u = self.net(torch.stack((x, y), dim=-1).view(-1, 2)).view(*x.shape)
        bc = x * (L - x) * y * (L - y)
        if self.technique in ('FBC', 'OG'):
            return bc * u
        if self.technique == 'FN':
            nf = torch.ones_like(x)
            for k in range(1, self.nx):
                nf = nf * (x - k * L / self.nx)
            for k in range(1, self.ny):
                nf = nf * (y - k * L / self.ny)
            return bc * nf * u
        raise ValueError(f"Unknown technique: {self.technique}")


def _envelope(model, L):
    t = getattr(model, "technique", None)
    if t in ('FBC', 'OG'):
        return poly_envelope(L)
    if t == 'FN':
        return poly_envelope(L, [[k * L / model.nx for k in range(1, model.nx)], [k * L / model.ny for k in range(1, model.ny)]])
    raise ValueError(f"Unknown technique: {t}")


def _points(x, y):
    return torch.stack((x.detach().reshape(-1), y.detach().reshape(-1)), dim=1).contiguous()


def k_squared(nx, ny, L):
    """2 m E / hbar^2 with E = ((nx pi)^2 + (ny pi)^2) / (2 L^2)   (IPW_2D.py:188-190)."""
    return (nx * np.pi) ** 2 / L ** 2 + (ny * np.pi) ** 2 / L ** 2


def PINN_loss(model, x, y, nx, ny, L=2.0):
    """mean((u_xx + u_yy + k^2 u)^2): the inline residual block of train_pinn_seperate (IPW_2D.py:195-224)."""
    return residual_mean(model, _points(x, y), ProgramSpec(_lib.PROG_PINN, alpha=1.0, beta_const=k_squared(nx, ny, L)),
                         _envelope(model, L))


def DRM_loss(model, x, y, L=2.0):
    """mean(|grad u|^2) / mean(u^2 + 1e-8)   (IPW_2D.py:225-228)."""
    m = residual_means(model, _points(x, y), ProgramSpec(_lib.PROG_RAYLEIGH, alpha=1.0), _envelope(model, L))
    return m[0] / (m[1] + 1e-8)


def data_loss(model, X_data, Y_data, u_data, L=2.0):
    """mean((u(X_data, Y_data) - u_data)^2)   (IPW_2D.py:230-232)."""
    return residual_mean(model, _points(X_data, Y_data), ProgramSpec(_lib.PROG_MSE), _envelope(model, L),
                         f=u_data.detach().reshape(-1))


def orthogonal_loss(model, x, y, nx, ny, L):
    """Projections on the exact states of lower energy (IPW_2D.py:113-126)."""
    u = values(model, _points(x, y), _envelope(model, L)).view(*x.shape)
    xd, yd = x.detach(), y.detach()
    ortho = 0.0
    for i in range(1, max(nx, ny) + 1):
        for j in range(1, max(nx, ny) + 1):
            if i ** 2 + j ** 2 < nx ** 2 + ny ** 2:
                ue = Exact_solution(L, i, j, xd, yd)
                inner = torch.mean(u * ue) * L * L
                ortho = ortho + inner ** 2 / (torch.mean(ue ** 2) * L * L + 1e-8)
    return ortho
