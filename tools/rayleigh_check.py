"""Development aid (GPU): gradient error of the Rayleigh-quotient (Deep Ritz eigenvalue) losses on the tensor-core kernel
and on the generic fp32 kernel, against the reference's float64 goldens at the BASELINE config-4 shapes, next to the bar
the fixture itself justifies (tests/conftest.grads_bar: max(1e-5, 2 x the reference's own float32 error))."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import pde_b200 as pb
from conftest import grads_bar, grads_err, grads_from, load_golden, net_from
from pde_b200.schrodinger import ipw_2d as I, qho_2d as Q


def run(name, path):
    g = load_golden(name)
    L, nx, ny = float(g["L"]), int(g["nx"]), int(g["ny"])
    Ws, bs = net_from(g)
    layers = [Ws[0].shape[1]] + [W.shape[0] for W in Ws]
    qho = "qho2d" in name
    tech = "FN" if "_fn_" in name else "FBC"
    model = (Q if qho else I).FCN(layers, nx, ny, tech).double()
    lin = [m for m in model.net if isinstance(m, torch.nn.Linear)]
    with torch.no_grad():
        for l, W, b in zip(lin, Ws, bs):
            l.weight.copy_(torch.tensor(W)); l.bias.copy_(torch.tensor(b))
    model = model.to("cuda", torch.float32)
    g1 = torch.linspace(-L if qho else 0.0, L, int(g["grid_n"]), dtype=torch.float64)
    x, y = torch.meshgrid(g1, g1, indexing="ij")
    x, y = x.to("cuda", torch.float32), y.to("cuda", torch.float32)
    with pb.ops.kernel_path(path):
        loss = Q.DRM_loss(model, x, y, L) if qho else I.DRM_loss(model, x, y, L)
        loss.backward()
    got = ([l.weight.grad.double().cpu().numpy() for l in lin], [l.bias.grad.double().cpu().numpy() for l in lin])
    e, where = grads_err(got, grads_from(g, "drm_"))
    lerr = abs(float(loss) - float(g["drm_loss"])) / abs(float(g["drm_loss"]))
    return lerr, e, where, grads_bar(g, "drm_", torch.float32), pb.ops.last_kernel_path()


for name in ("cfg4_qho2d_fbc_00", "cfg4_qho2d_fn_21", "cfg4_ipw2d_fbc_11", "cfg4_ipw2d_fn_32"):
    for path in ("simt", "tc"):
        lerr, e, where, bar, used = run(name, path)
        print(f"{name:20s} {path:5s} ({used:8s}) loss rel {lerr:.2e}  grad err {e:.2e} ({where})  bar {bar:.2e}  {'OK' if e <= bar else 'ABOVE BAR'}", flush=True)
