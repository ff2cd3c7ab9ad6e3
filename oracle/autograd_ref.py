"""TEST INFRASTRUCTURE ONLY — the reference's *algorithm* (nested torch autograd on CPU).

``/root/reference`` is a set of Python scripts that cannot travel to the GPU box, so
this file restates, in torch, the exact operator sequence the reference executes for
the collocation-point loss step: build the MLP, call ``autograd.grad(create_graph=True)``
once for ∇u and once more per dimension for the Hessian diagonal, reduce to a scalar
and call ``.backward()``.  It is used (a) as a second checker next to
``oracle/jets_numpy.py`` and (b) as the thing ``bench.py`` times for ``cpu_baseline`` /
``--impl reference`` (kind "port": same ops, same autograd graph, same thread pool as the
reference's CPU path).  Parity status: PINNED by ``tests/golden/*.npz`` (generated from
the live reference by ``tests/golden/make_golden.py``); see ``tests/test_oracle.py``.

Sites followed: Poisson_Equations/Poisson_ND.py:11-33 (network + hard-BC envelope),
:61-71 (gradient / Laplacian helpers), :91-103 (PINN and Deep-Ritz losses),
:74-88 and :105-128 (bump weight and WAN losses).
"""
from __future__ import annotations

import torch
import torch.nn as nn


class _Sine(nn.Module):
    def forward(self, t):
        return t.sin()


def build_mlp(widths, activation="sin", dtype=torch.float32):
    """Linear/activation stack with the reference's layer ordering (Linear at even
    indices of an ``nn.Sequential``), default ``nn.Linear`` initialisation."""
    act = _Sine if activation == "sin" else nn.Tanh
    mods = []
    for i in range(len(widths) - 2):
        mods.append(nn.Linear(widths[i], widths[i + 1]))
        mods.append(act())
    mods.append(nn.Linear(widths[-2], widths[-1]))
    return nn.Sequential(*mods).to(dtype)


def load_params(net, Ws, bs):
    lin = [m for m in net if isinstance(m, nn.Linear)]
    with torch.no_grad():
        for m, W, b in zip(lin, Ws, bs):
            m.weight.copy_(torch.as_tensor(W, dtype=m.weight.dtype))
            m.bias.copy_(torch.as_tensor(b, dtype=m.bias.dtype))


def solution(net, X, L, bc_mode):
    """Network output with the optional hard-Dirichlet envelope Π x_i (L − x_i)."""
    out = net(X)
    if bc_mode == "FBC":
        return (X * (L - X)).prod(dim=1, keepdim=True) * out
    if bc_mode == "RB":
        return out
    raise ValueError("bc_mode must be 'FBC' or 'RB'")


def _grad(y, X):
    return torch.autograd.grad(y, X, torch.ones_like(y), create_graph=True)[0]


def _laplace(u, X):
    g = _grad(u, X)
    total = None
    for i in range(X.shape[1]):
        col = g[:, i:i + 1]
        hii = _grad(col, X)[:, i:i + 1]
        total = hii if total is None else total + hii
    return total


def pinn_loss(net, X, f, L, bc_mode="FBC"):
    u = solution(net, X, L, bc_mode)
    r = -_laplace(u, X) - f
    return (r * r).mean()


def drm_loss(net, X, f, L, bc_mode="FBC"):
    u = solution(net, X, L, bc_mode)
    g = _grad(u, X)
    return (0.5 * (g * g).sum(dim=1, keepdim=True) - f * u).mean()


def bump(X, L):
    half = L / 2.0
    t = (X - half) / half
    inside = (t.abs() < 1.0).to(X.dtype)
    phi = torch.exp(1.0 / (t * t - 1.0)) / 0.210987 * inside
    w = phi.prod(dim=1, keepdim=True)
    dw = torch.nan_to_num(_grad(w, X))
    return w, dw


def wan_losses(u_net, v_net, X, f, L, bc_mode="FBC", eps=1e-8, v_reg_weight=0.0):
    u = solution(u_net, X, L, bc_mode)
    v = v_net(X)
    w, dw = bump(X, L)
    gu, gv = _grad(u, X), _grad(v, X)
    phi = w * v
    gphi = dw * v + w * gv
    weak = ((gu * gphi).sum(dim=1, keepdim=True) - f * phi).mean()
    pn = (phi * phi).mean()
    loss_u = weak * weak / (pn + eps)
    reg = ((gv * gv).sum(dim=1, keepdim=True) + v * v).mean()
    loss_v = -torch.log(loss_u + eps) + v_reg_weight * reg
    return loss_u, loss_v, weak.detach(), pn.detach()


def manufactured_rhs(X, L, ks):
    """f = (Σ (k_i π / L)²) Π sin(k_i π x_i / L)   (Poisson_ND.py:49-58)."""
    import math
    u = torch.ones(X.shape[0], 1, dtype=X.dtype, device=X.device)
    s = 0.0
    for i, k in enumerate(ks):
        u = u * torch.sin(k * math.pi * X[:, i:i + 1] / L)
        s += (k * math.pi / L) ** 2
    return s * u


def loss_and_grads(kind, net, X, f, L, bc_mode="FBC", chunk=None):
    """Loss value + flat parameter gradient the way a training step would get them.

    ``chunk``: evaluate in point chunks and accumulate ``p.grad`` (exact for the
    plain-mean Poisson losses); this is how a 2^22-point batch is timed on CPU,
    where a single autograd graph of that size does not fit in host memory.
    """
    fn = pinn_loss if kind == "pinn" else drm_loss
    for p in net.parameters():
        p.grad = None
    N = X.shape[0]
    chunk = N if chunk is None else chunk
    total = 0.0
    for s in range(0, N, chunk):
        Xc = X[s:s + chunk].detach().clone().requires_grad_(True)
        fc = f[s:s + chunk]
        part = fn(net, Xc, fc, L, bc_mode) * (Xc.shape[0] / N)
        part.backward()
        total += float(part.detach())
    g = torch.cat([(torch.zeros_like(p) if p.grad is None else p.grad).reshape(-1) for p in net.parameters()])
    return total, g
