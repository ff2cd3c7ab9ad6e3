"""Drop-in for Schrodinger_Equations/Quantum_Harmonic_Oscillator/QHO_2D_Energy.py: QHO_2D.py with a
trainable energy in the PINN residual (``E_train``, QHO_2D_Energy.py:287-291,382-383).  The network,
envelopes and the DRM / WAN losses are those of ``qho_2d``; ``PINN_loss`` takes the energy as a
scalar Parameter and its ``.backward()`` fills ``E_train.grad`` from the same fused launch."""
import torch
import torch.nn as nn

from .qho_2d import DRM_loss, Exact_energy, FCN, PINN_loss, WAN_loss, hermite_nodes  # noqa: F401


def make_trainable_energy(nx, ny, L, device=None):
    """E_train = nn.Parameter(Exact_energy(nx, ny, L))   (QHO_2D_Energy.py:288)."""
    return nn.Parameter(torch.tensor(Exact_energy(nx, ny, L), device=device))
