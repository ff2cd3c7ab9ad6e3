"""Drop-in for Schrodinger_Equations/Infinite_Potential_Well/IPW_1D_WAN.py (weak adversarial network)."""
import numpy as np
import torch
import torch.nn as nn

from ..ops import WanSpec, wan_means, wan_scalar_losses
from ._common import NO_ENVELOPE, mlp, poly_envelope


def Exact_energy(n, L):
    """E_n = (n pi)^2 / (2 L^2)   (IPW_1D_WAN.py:25-29)."""
    return (n * np.pi) ** 2 / (2 * L ** 2)


class FCN(nn.Module):
    """tanh network, optional x (L - x) envelope (IPW_1D_WAN.py:62-86)."""

    def __init__(self, layers, num_states=1, L=2.0, enforce_bc=False):
        super().__init__()
        self.enforce_bc, self.num_states, self.L = enforce_bc, num_states, L
        self.net = mlp(layers, nn.Tanh)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight, gain=nn.init.calculate_gain('tanh'))
                nn.init.zeros_(m.bias)

    def forward(self, x):
        y = self.net(x)
        return x * (self.L - x) * y if self.enforce_bc else y


def _envelope(model):
    return poly_envelope(getattr(model, "L", 2.0)) if getattr(model, "enforce_bc", False) else NO_ENVELOPE


def WAN_loss(u_model, v_model, x, n, L, weight_pde=1.0, weight_norm=1.0, *, u_jets=None, v_jets=None):
    """(total_loss, loss_v, loss_pde, loss_norm)   (IPW_1D_WAN.py:88-115): weak residual of
    -1/2 u'' = E_n u against phi = w v, normalised by mean(phi^2), plus (L mean(u^2) - 1)^2."""
    m = wan_means(u_model, v_model, x, WanSpec(alpha=0.5, energy_const=Exact_energy(n, L), w_lo=0.0, w_hi=float(L)),
                  env_u=_envelope(u_model), env_v=_envelope(v_model), u_jets=u_jets, v_jets=v_jets)
    loss_pde, loss_v, loss_norm, total_loss = wan_scalar_losses(m, kind=0, vol=L, w_pde=weight_pde, w_norm=weight_norm)
    return total_loss, loss_v, loss_pde, loss_norm
