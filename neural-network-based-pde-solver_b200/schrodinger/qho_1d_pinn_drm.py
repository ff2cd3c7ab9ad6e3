"""Drop-in for Schrodinger_Equations/Quantum_Harmonic_Oscillator/QHO_1D_PINN_DRM.py (1-D oscillator,
omega = sqrt 2, sine network held in a ModuleList)."""
import math

import numpy as np
import torch
import torch.nn as nn

from .. import _lib
from ..ops import ProgramSpec, residual_mean, residual_means
from ._common import NO_ENVELOPE, values, window_envelope
from .qho_2d import hermite_nodes

OMEGA = math.sqrt(2)


def phys_hermite(n, x):
    """Physicists' Hermite polynomial by recurrence (QHO_1D_PINN_DRM.py:25-39)."""
    if n == 0:
        return torch.ones_like(x)
    if n == 1:
        return 2 * x
    a, b = torch.ones_like(x), 2 * x
    for k in range(2, n + 1):
        a, b = b, 2 * x * b - 2 * (k - 1) * a
    return b


def Exact_solution(n, x, omega=OMEGA):
    """(QHO_1D_PINN_DRM.py:40-46)."""
    Hn = phys_hermite(n, torch.sqrt(torch.tensor(omega)) * x)
    norm = (omega / np.pi) ** 0.25 / math.sqrt(2 ** n * math.factorial(n))
    return norm * Hn * torch.exp(-omega * x ** 2 / 2)


def Potential(x, omega=OMEGA):
    """1/2 omega^2 x^2   (QHO_1D_PINN_DRM.py:48-49)."""
    return 0.5 * omega ** 2 * x ** 2


def Energy(n, omega=OMEGA):
    """(n + 1/2) omega   (QHO_1D_PINN_DRM.py:51-53)."""
    return (n + 0.5) * omega


class SineActivation(nn.Module):
    def forward(self, x):
        return torch.sin(x)


class FCN(nn.Module):
    """Linear layers in a ModuleList, sine between them (QHO_1D_PINN_DRM.py:57-72)."""

    def __init__(self, layers):
        super().__init__()
        self.activation = SineActivation()
        self.layers = nn.ModuleList(nn.Linear(layers[i], layers[i + 1]) for i in range(len(layers) - 1))

    def forward(self, x):
        for layer in self.layers[:-1]:
            x = self.activation(layer(x))
        return self.layers[-1](x)


class FCN_Single(nn.Module):
    """Trial function: exp-window boundary factor and / or forced nodes (QHO_1D_PINN_DRM.py:97-154)."""

    def __init__(self, layers, num_states=1, domain_length=20.0, enforce_bc=False, FN=False):
        super().__init__()
        self.net = FCN(layers)
        self.num_states, self.domain_length = num_states, domain_length
        self.enforce_bc, self.FN = enforce_bc, FN
        self.energies = nn.Parameter(torch.tensor(Energy(num_states), dtype=torch.float32))
        self.nodes = {n: hermite_nodes(n) for n in range(1, 6)}

    def forward(self, x):
        L = self.domain_length / 2.0
        out = self.net(x)
        win = (1 - torch.exp(-(x + L))) * (1 - torch.exp(x - L))
        if self.FN and self.num_states in self.nodes:
            trial = torch.ones_like(x)
            for node in self.nodes[self.num_states].to(x.device):
                trial = trial * (x - node)
            if self.enforce_bc:
                trial = trial * win
            return out * trial
        if self.enforce_bc:
            return out * win
        return out


def _envelope(model):
    L = float(model.domain_length) / 2.0
    nodes = None
    if getattr(model, "FN", False) and model.num_states in model.nodes:
        nodes = [[float(v) for v in model.nodes[model.num_states]]]
        if not model.enforce_bc:
            from ..ops import EnvelopeSpec
            return EnvelopeSpec(_lib.ENV_NONE, 0.0, 0.0, nodes)
        return window_envelope(L, nodes)
    return window_envelope(L) if getattr(model, "enforce_bc", False) else NO_ENVELOPE


def PINN_loss(model, x):
    """mean((-1/2 u'' + V u - E_n u)^2)   (QHO_1D_PINN_DRM.py:161-174)."""
    V = Potential(x.detach())
    return residual_mean(model, x, ProgramSpec(_lib.PROG_PINN, alpha=-0.5, energy_const=Energy(model.num_states)),
                         _envelope(model), beta=V)


def DRM_loss(model, x):
    """mean(1/2 u'^2 + V u^2) / mean(u^2)   (QHO_1D_PINN_DRM.py:176-185)."""
    V = Potential(x.detach())
    m = residual_means(model, x, ProgramSpec(_lib.PROG_RAYLEIGH, alpha=0.5), _envelope(model), beta=V)
    return m[0] / m[1]


def normalization_loss(model, x):
    """(sqrt(sum(u^2) dx) - 1)^2 on a uniform grid (QHO_1D_PINN_DRM.py:187-195)."""
    m2 = residual_mean(model, x, ProgramSpec(_lib.PROG_MSE), _envelope(model))
    xd = x.detach()
    dx = (xd[1] - xd[0]).reshape(())
    return (torch.sqrt(m2 * xd.shape[0] * dx) - 1) ** 2


def Orthogonal_loss(model, x, n, domain_length):
    """sum_k <u, u_k>^2 / <u_k, u_k> over the lower exact states (QHO_1D_PINN_DRM.py:197-212)."""
    if n == 0:
        return torch.tensor(0.0, device=x.device)
    u = values(model, x, _envelope(model))
    out = 0.0
    for k in range(n):
        ue = Exact_solution(k, x.detach())
        inner = torch.mean(u * ue) * 2 * domain_length
        out = out + inner ** 2 / (torch.mean(ue ** 2) * 2 * domain_length)
    return out
