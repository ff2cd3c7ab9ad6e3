"""Drop-in for Schrodinger_Equations/Kramers_Henneberger/KH_1D.py (KH-frame 1-D eigenproblem,
trainable energy)."""
import torch
import torch.nn as nn

from .. import _lib
from ..ops import NO_ENVELOPE, ProgramSpec, WanSpec, residual_mean, residual_means, wan_means, wan_scalar_losses
from ._common import Sin, mlp, window_envelope


def V_base(x, V0=-24.856):
    """V0 exp(-sqrt(x^2 + 16)) / sqrt(x^2 + 6.27^2)   (KH_1D.py:23-24)."""
    return V0 * torch.exp(-(x ** 2 + 16.0).sqrt()) / (x ** 2 + 6.27 ** 2).sqrt()


def V_KH(x, alpha=0.0, V0=-24.856, use_avg=True, n_theta=500):
    """Shifted or cycle-averaged KH potential (KH_1D.py:27-43)."""
    if not use_avg:
        return V_base(x + alpha, V0)
    if alpha == 0.0:
        return V_base(x, V0)
    # the reference's theta grid is float32 (default dtype); its sine is taken on the host so that the
    # potential is bit-identical on every device (CPU and GPU sinf differ in the last ulp)
    sin_th = torch.sin(torch.linspace(0, 2 * torch.pi, n_theta)).to(x.device)
    return V_base(x[..., None] + alpha * sin_th[None, ...], V0).mean(dim=-1)


_V_CACHE = []   # [(weakref to the grid tensor, its _version, (alpha, V0, use_avg, n_theta), V)]


def _potential(x, alpha, V0, use_avg, n_theta):
    """The potential depends on neither the parameters nor the epoch; the reference rebuilds an
    (N, n_theta) tensor in every loss call (KH_1D.py:231,239,259) — here it is computed once per grid.
    An entry is reused only for the very same tensor object at the same version: storage addresses are
    recycled by the caching allocator, so they cannot identify a grid."""
    import weakref
    cfg = (float(alpha), float(V0), bool(use_avg), int(n_theta))
    alive = []
    hit = None
    for ref, ver, c, V in _V_CACHE:
        t = ref()
        if t is None:
            continue
        alive.append((ref, ver, c, V))
        if t is x and ver == x._version and c == cfg:
            hit = V
    _V_CACHE[:] = alive[-16:]
    if hit is not None:
        return hit
    with torch.no_grad():
        V = V_KH(x.detach(), alpha=alpha, V0=V0, use_avg=use_avg, n_theta=n_theta)
    _V_CACHE.append((weakref.ref(x), x._version, cfg, V))
    return V


class FCN1D(nn.Module):
    """sin network, 'RAW' or exp-window 'FBC' output (KH_1D.py:104-124)."""

    def __init__(self, layers, technique='RAW'):
        super().__init__()
        self.technique = technique
        self.net = mlp(layers, Sin)

    def forward(self, x, L=10.0):
        u = self.net(x.view(-1, 1)).view_as(x)
        if self.technique == 'FBC':
            return (1 - torch.exp(-(x + L))) * (1 - torch.exp(x - L)) * u
        if self.technique == 'RAW':
            return u
        raise ValueError(f"Unknown technique {self.technique}")


class UnifiedEigenModel(nn.Module):
    """u-network plus trainable energy (KH_1D.py:214-223)."""

    def __init__(self, layers=[1, 64, 64, 64, 1], technique='RAW', E_init=0.0, device=None):
        super().__init__()
        self.u_model = FCN1D(layers, technique=technique)
        self.energy = nn.Parameter(torch.tensor(float(E_init)))
        if device is not None:
            self.to(device)

    def forward(self, x, L=10.0):
        return self.u_model(x, L=L)


def _envelope(net, L):
    t = getattr(getattr(net, "u_model", net), "technique", 'RAW')
    if t == 'FBC':
        return window_envelope(L)
    if t == 'RAW':
        return NO_ENVELOPE
    raise ValueError(f"Unknown technique {t}")


def pinn_loss(model, x, alpha, V0, use_avg=True, n_theta=500):
    """mean((-1/2 u'' + V u - E u)^2) with the envelope half-width max|x| (KH_1D.py:226-234)."""
    env = NO_ENVELOPE
    if getattr(model.u_model, "technique", 'RAW') != 'RAW':
        env = _envelope(model, x.detach().abs().max().item())      # the reference syncs here too (:227)
    Vx = _potential(x, alpha, V0, use_avg, n_theta)
    return residual_mean(model, x, ProgramSpec(_lib.PROG_PINN, alpha=-0.5), env, beta=Vx, energy=model.energy)


def drm_loss(model, x, alpha, V0, L, use_avg=True, n_theta=500):
    """2L mean(1/2 u'^2 + V u^2) / (2L mean(u^2) + 1e-12)   (KH_1D.py:236-242)."""
    Vx = _potential(x, alpha, V0, use_avg, n_theta)
    m = residual_means(model, x, ProgramSpec(_lib.PROG_RAYLEIGH, alpha=0.5), _envelope(model, L), beta=Vx)
    return (2 * L) * m[0] / ((2 * L) * m[1] + 1e-12)


def wan_loss(model, v_model, x, alpha, V0, L, use_avg=True, n_theta=500, *, u_jets=None, v_jets=None):
    """(pde_loss, norm_u)   (KH_1D.py:244-269): (I_full / ||phi||^2)^2 and (int u^2 - 1)^2."""
    Vx = _potential(x, alpha, V0, use_avg, n_theta)
    m = wan_means(model, v_model, x, WanSpec(alpha=0.5, w_lo=-float(L), w_hi=float(L), eps_den=1e-10),
                  env_u=_envelope(model, L), env_v=_envelope(v_model, L), beta=Vx, energy=model.energy,
                  u_jets=u_jets, v_jets=v_jets)
    pde_loss, _, norm_u, _ = wan_scalar_losses(m, kind=1, eps_pde=1e-12, vol=2 * L)
    return pde_loss, norm_u
