"""Drop-in for Schrodinger_Equations/Infinite_Potential_Well/IPW_1D_PINN_DRM.py (1-D infinite well)."""
import math

import numpy as np
import torch
import torch.nn as nn

from .. import _lib
from ..ops import ProgramSpec, residual_mean, residual_means
from ._common import NO_ENVELOPE, mlp, poly_envelope


class FCN(nn.Module):
    """tanh network with optional hard boundary / forced-node ansatz (IPW_1D_PINN_DRM.py:32-61)."""

    def __init__(self, layers, num_states=1, L=2.0, enforce_bc=False, FN=False):
        super().__init__()
        self.enforce_bc, self.FN, self.num_states, self.L = enforce_bc, FN, num_states, L
        self.net = mlp(layers, nn.Tanh)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight, gain=nn.init.calculate_gain('tanh'))
                nn.init.zeros_(m.bias)
        # node positions k L / n, stored in float32 like the reference (:38-40)
        self.nodes = {n: torch.tensor([k * L / n for k in range(1, n)], dtype=torch.float32) for n in range(1, 11)}

    def forward(self, x):
        y = self.net(x)
        if self.FN and self.num_states in self.nodes:
            trial = x * (self.L - x)
            for node in self.nodes[self.num_states].to(x.device):
                trial = trial * (x - node)
            return y * trial
        if self.enforce_bc:
            return y * (x * (self.L - x))
        return y


def _envelope(model):
    L = float(getattr(model, "L", 2.0))
    if getattr(model, "FN", False) and model.num_states in model.nodes:
        return poly_envelope(L, [[float(v) for v in model.nodes[model.num_states]]])
    if getattr(model, "enforce_bc", False):
        return poly_envelope(L)
    return NO_ENVELOPE


def PINN_loss(model, x, n, L):
    """mean((u'' + k^2 u)^2), k^2 = (n pi / L)^2   (IPW_1D_PINN_DRM.py:63-83)."""
    E = (n * np.pi) ** 2 / (2 * L ** 2)
    return residual_mean(model, x, ProgramSpec(_lib.PROG_PINN, alpha=1.0, beta_const=2.0 * E), _envelope(model))


def DRM_loss(model, x):
    """Rayleigh quotient mean(u'^2) / mean(u^2)   (IPW_1D_PINN_DRM.py:85-90)."""
    m = residual_means(model, x, ProgramSpec(_lib.PROG_RAYLEIGH, alpha=1.0), _envelope(model))
    return m[0] / m[1]
