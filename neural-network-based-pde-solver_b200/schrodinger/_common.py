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


__all__ = ["mlp", "poly_envelope", "window_envelope", "NO_ENVELOPE", "Sin", "torch", "nn"]
