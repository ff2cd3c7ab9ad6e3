"""Drop-in for Schrodinger_Equations/Infinite_Potential_Well/IPW_1D_WAN_FN.py (weak adversarial
network with the forced-node ansatz)."""
import torch
import torch.nn as nn

from ..ops import WanSpec, wan_means, wan_scalar_losses
from ._common import mlp, poly_envelope
from .ipw_1d_wan import Exact_energy  # noqa: F401  (IPW_1D_WAN_FN.py:24-27)


class FCN(nn.Module):
    """tanh network times x (L - x) prod_{j<n} (x - j L / n), n = num_states (IPW_1D_WAN_FN.py:60-88;
    the loop over k keeps only its last factor)."""

    def __init__(self, layers, num_states=1, L=2.0, enforce_bc=False):
        super().__init__()
        self.enforce_bc, self.num_states, self.L = enforce_bc, num_states, L
        self.net = mlp(layers, nn.Tanh)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight, gain=nn.init.calculate_gain('tanh'))
                nn.init.zeros_(m.bias)

    def forward(self, x):
        L, n = self.L, self.num_states
        f = x * (L - x)
        for j in range(1, n):
            f = f * (x - j * L / n)
        return f * self.net(x)


def _envelope(model):
    L, n = float(model.L), int(model.num_states)
    return poly_envelope(L, [[j * L / n for j in range(1, n)]])


def WAN_loss(u_model, v_model, x, n, L, weight_pde=1.0, weight_norm=1.0, *, u_jets=None, v_jets=None):
    """(total_loss, loss_v, loss_pde, loss_norm)   (IPW_1D_WAN_FN.py:91-118)."""
    m = wan_means(u_model, v_model, x, WanSpec(alpha=0.5, energy_const=Exact_energy(n, L), w_lo=0.0, w_hi=float(L)),
                  env_u=_envelope(u_model), env_v=_envelope(v_model), u_jets=u_jets, v_jets=v_jets)
    loss_pde, loss_v, loss_norm, total_loss = wan_scalar_losses(m, kind=0, vol=L, w_pde=weight_pde, w_norm=weight_norm)
    return total_loss, loss_v, loss_pde, loss_norm
