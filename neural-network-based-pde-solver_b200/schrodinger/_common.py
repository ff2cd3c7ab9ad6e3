"""Shared pieces of the Schrödinger drop-ins."""
import torch
import torch.nn as nn

from .. import _lib
from ..ops import NO_ENVELOPE, EnvelopeSpec
from ..poisson import Sin


def mlp(layers, act):
    mods = []
    for i in range(len(layers) - 2):
        mods += [nn.Linear(layers[i], layers[i + 1]), act()]
    mods.append(nn.Linear(layers[-2], layers[-1]))
    return nn.Sequential(*mods)


def poly_envelope(L, nodes=None):
    return EnvelopeSpec(_lib.ENV_POLY, 0.0, float(L), nodes or [])


def window_envelope(L, nodes=None):
    return EnvelopeSpec(_lib.ENV_EXPWIN, -float(L), float(L), nodes or [])


def envelope_values(env, X):
    """B(x) = prod_i b_i(x_i) of an EnvelopeSpec evaluated with torch ops (value-only terms)."""
    B = torch.ones(X.shape[0], 1, dtype=X.dtype, device=X.device)
    for i in range(X.shape[1]):
        t = X[:, i:i + 1]
        if env.kind == _lib.ENV_POLY:
            B = B * (t - env.lo) * (env.hi - t)
        elif env.kind == _lib.ENV_EXPWIN:
            B = B * (1 - torch.exp(-(t - env.lo))) * (1 - torch.exp(t - env.hi))
        if i < len(env.nodes):
            for node in env.nodes[i]:
                B = B * (t - float(node))
    return B


def values(model, X, env=NO_ENVELOPE):
    """u(X) = B(X) N(X) with the network evaluated by the jet kernel at order 0 (value-only terms:
    normalisation, orthogonality, symmetry, data — QHO_1D_PINN_DRM.py:185-212, IPW_2D.py:113-126),
    differentiable with respect to the parameters."""
    from ..ops import mlp_jets
    Xd = X.detach()
    if Xd.dim() == 1:
        Xd = Xd.view(-1, 1)
    return envelope_values(env, Xd) * mlp_jets(model, Xd, 0)[:, 0:1]


__all__ = ["mlp", "poly_envelope", "window_envelope", "envelope_values", "values", "NO_ENVELOPE", "Sin", "torch", "nn"]
