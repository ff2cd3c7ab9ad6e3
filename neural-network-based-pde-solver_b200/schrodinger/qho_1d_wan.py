"""Drop-in for Schrodinger_Equations/Quantum_Harmonic_Oscillator/QHO_1D_WAN.py (weak adversarial
network for the 1-D oscillator, trainable energy)."""
import torch
import torch.nn as nn

from ..ops import WanSpec, wan_means, wan_scalar_losses
from ._common import NO_ENVELOPE, mlp, window_envelope
from .qho_1d_pinn_drm import Energy, Exact_solution, Potential, phys_hermite  # noqa: F401  (same helpers, :25-53)


class FCN(nn.Module):
    """tanh network, optional exp-window envelope, trainable ``energies`` (QHO_1D_WAN.py:86-113)."""

    def __init__(self, layers, num_states=1, L=10.0, enforce_bc=False):
        super().__init__()
        self.enforce_bc, self.num_states, self.L = enforce_bc, num_states, L
        self.net = mlp(layers, nn.Tanh)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight, gain=nn.init.calculate_gain('tanh'))
                nn.init.zeros_(m.bias)
        self.energies = nn.Parameter(torch.tensor(Energy(num_states), dtype=torch.float32))

    def forward(self, x):
        y = self.net(x)
        if self.enforce_bc:
            return y * (1 - torch.exp(-(x + self.L))) * (1 - torch.exp(x - self.L))
        return y


def _envelope(model):
    return window_envelope(float(model.L)) if getattr(model, "enforce_bc", False) else NO_ENVELOPE


def WAN_loss(u_model, v_model, x, n, L, weight_pde=1.0, weight_norm=1.0, *, u_jets=None, v_jets=None):
    """(total_loss, loss_v, loss_pde, loss_norm)   (QHO_1D_WAN.py:115-140): weak residual of
    -1/2 u'' + V u = E u against phi = w v with E = ``u_model.energies`` (trainable),
    plus (2 L mean(u^2) - 1)^2."""
    V = Potential(x.detach())
    m = wan_means(u_model, v_model, x, WanSpec(alpha=0.5, w_lo=-float(L), w_hi=float(L)),
                  env_u=_envelope(u_model), env_v=_envelope(v_model), beta=V, energy=u_model.energies, u_jets=u_jets, v_jets=v_jets)
    loss_pde, loss_v, loss_norm, total_loss = wan_scalar_losses(m, kind=0, vol=2 * L, w_pde=weight_pde, w_norm=weight_norm)
    return total_loss, loss_v, loss_pde, loss_norm
