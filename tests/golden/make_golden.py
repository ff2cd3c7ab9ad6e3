"""Generate the golden fixtures in this directory from the LIVE reference.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
The reference holds no golden vectors of its own, so these fixtures —
outputs of the reference's own loss functions and ``.backward()`` on seeded weights and
points, in float64 — are what pins ``oracle/`` and, through it, the CUDA path.
Nothing in ``tests/`` reads /root/reference at run time; only this script does.
"""
from __future__ import annotations

import importlib.util
import math
import os
import sys
import tempfile
import types

import numpy as np
import torch

REF = os.environ.get("PDE_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def load_ref(relpath, name):
    """Import a reference script by path: stub matplotlib, run from a temp cwd
    (the Schrödinger scripts create results/ folders at import time)."""
    for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.cm"):
        if m not in sys.modules:
            stub = types.ModuleType(m)
            sys.modules[m] = stub
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].use = lambda *a, **k: None
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp()
    os.chdir(tmp)
    try:
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        os.chdir(cwd)
    return mod


def linears(seq):
    return [m for m in seq if isinstance(m, torch.nn.Linear)]


def params_np(seq, prefix):
    d = {}
    for i, m in enumerate(linears(seq)):
        d[f"{prefix}W{i}"] = m.weight.detach().numpy().copy()
        d[f"{prefix}b{i}"] = m.bias.detach().numpy().copy()
    return d


def grads_np(seq, prefix):
    d = {}
    for i, m in enumerate(linears(seq)):
        gw = m.weight.grad
        gb = m.bias.grad
        d[f"{prefix}gW{i}"] = (torch.zeros_like(m.weight) if gw is None else gw).detach().numpy().copy()
        d[f"{prefix}gb{i}"] = (torch.zeros_like(m.bias) if gb is None else gb).detach().numpy().copy()
    return d


def zero_grads(*mods):
    for m in mods:
        for p in m.parameters():
            p.grad = None


def save(name, **arrs):
    path = os.path.join(OUT, name + ".npz")
    np.savez(path, **arrs)
    print(f"wrote {name}.npz  ({os.path.getsize(path) / 1024:.0f} KiB)")


# ------------------------------------------------------------------ Poisson
def poisson_cases(P):
    L = 2.0
    # (tag, dim, width, depth, bc_mode, method, N)
    cases = [
        ("poisson_pinn_d1_w64_fbc", 1, 64, 5, "FBC", "pinn", 64),   # BASELINE config 1 shape
        ("poisson_pinn_d3_w64_fbc", 3, 64, 5, "FBC", "pinn", 64),   # BASELINE config 2 shape
        ("poisson_drm_d5_w64_rb", 5, 64, 5, "RB", "drm", 64),       # BASELINE config 3 shape
        ("poisson_pinn_d2_w16_fbc", 2, 16, 4, "FBC", "pinn", 96),
        ("poisson_pinn_d5_w16_fbc", 5, 16, 3, "FBC", "pinn", 96),
        ("poisson_pinn_d3_w16_rb", 3, 16, 4, "RB", "pinn", 96),
        ("poisson_pinn_d4_w12_fbc", 4, 12, 3, "FBC", "pinn", 50),   # ragged width / N
        ("poisson_drm_d1_w16_fbc", 1, 16, 4, "FBC", "drm", 96),
        ("poisson_drm_d2_w16_fbc", 2, 16, 4, "FBC", "drm", 96),
        ("poisson_drm_d3_w16_fbc", 3, 16, 4, "FBC", "drm", 96),
    ]
    for k, (tag, d, w, depth, bc, method, N) in enumerate(cases):
        torch.manual_seed(100 + k)
        model = P.SolutionNet(d, w, depth, bc_mode=bc).double()
        X = (torch.rand(N, d, dtype=torch.float64) * L).requires_grad_(True)
        ks = [1 + (i % 2) for i in range(d)]
        f = P.rhs_f_for_u_sin(X, L, ks).detach()
        fn = P.pinn_residual_loss if method == "pinn" else P.drm_energy_loss
        zero_grads(model)
        loss = fn(model, X, f, L)
        loss.backward()
        # also the raw network jets the reference's helpers produce (u, grad u, laplacian)
        X2 = X.detach().clone().requires_grad_(True)
        u = model(X2, L)
        g = P.grad_scalar_field(u, X2)
        lap = P.laplacian(u, X2)
        save(tag, X=X.detach().numpy(), f=f.numpy(), L=np.float64(L), loss=np.float64(loss.item()),
             u=u.detach().numpy(), grad_u=g.detach().numpy(), lap_u=lap.detach().numpy(),
             **params_np(model.net, ""), **grads_np(model.net, ""))

    # survey sanity pins: default dtype float64, seed 0, model first, then X = rand(512,d)*2
    pins = {}
    torch.set_default_dtype(torch.float64)
    for d in (1, 2, 3, 5):
        torch.manual_seed(0)
        m = P.SolutionNet(d, 64, 5, "FBC").double()
        X = (torch.rand(512, d) * 2.0).requires_grad_(True)
        f = P.rhs_f_for_u_sin(X, 2.0, [1] * d).detach()
        pins[f"pinn_d{d}"] = P.pinn_residual_loss(m, X, f, 2.0).item()
        pins[f"drm_d{d}"] = P.drm_energy_loss(m, X, f, 2.0).item()
    expect = {"pinn_d1": 3.997824334710243, "pinn_d2": 6.755232513291217, "pinn_d3": 6.414617763140292,
              "pinn_d5": 4.878397382574168, "drm_d1": 0.2063688668551195, "drm_d2": 0.0828105207704522,
              "drm_d3": -0.04937656836661489, "drm_d5": 1.869090502863647e-4}
    torch.set_default_dtype(torch.float32)
    for k, v in expect.items():
        assert abs(pins[k] - v) <= 1e-12 * max(1.0, abs(v)), (k, pins[k], v)
    print("survey sanity pins reproduced:", {k: float(f"{v:.6g}") for k, v in pins.items()})

    # WAN (two networks), d=2
    torch.manual_seed(200)
    um = P.SolutionNet(2, 16, 4, bc_mode="FBC").double()
    vm = P.CriticNet(2, 16, 3).double()
    N = 96
    X = (torch.rand(N, 2, dtype=torch.float64) * L).requires_grad_(True)
    f = P.rhs_f_for_u_sin(X, L, [1, 2]).detach()
    out = {}
    zero_grads(um, vm)
    lu, lv, weak, pn = P.wan_losses(um, vm, X, f, L, v_reg_weight=1.0)
    lu.backward(retain_graph=True)
    out.update({"lu_" + k: v for k, v in grads_np(um.net, "u_").items()})
    out.update({"lu_" + k: v for k, v in grads_np(vm.net, "v_").items()})
    zero_grads(um, vm)
    lv.backward()
    out.update({"lv_" + k: v for k, v in grads_np(um.net, "u_").items()})
    out.update({"lv_" + k: v for k, v in grads_np(vm.net, "v_").items()})
    save("poisson_wan_d2_w16", X=X.detach().numpy(), f=f.numpy(), L=np.float64(L), v_reg_weight=np.float64(1.0),
         loss_u=np.float64(lu.item()), loss_v=np.float64(lv.item()), weak=np.float64(weak.item()),
         phi_norm=np.float64(pn.item()), **params_np(um.net, "u_"), **params_np(vm.net, "v_"), **out)


# ------------------------------------------------------------------ 1-D well
def ipw_cases(I, W):
    L = 2.0
    for tag, kw, n in (("ipw1d_fbc_n2", dict(enforce_bc=True), 2), ("ipw1d_fn_n3", dict(FN=True, num_states=3), 3)):
        torch.manual_seed(300 + n)
        model = I.FCN([1, 20, 20, 1], L=L, **kw).double()
        with torch.no_grad():      # reference zero-initialises biases; perturb so bias paths are exercised
            for m in linears(model.net):
                m.bias.uniform_(-0.3, 0.3)
        x = torch.linspace(0.0, L, 65, dtype=torch.float64).view(-1, 1).requires_grad_(True)
        zero_grads(model)
        lp = I.PINN_loss(model, x, n, L)
        lp.backward()
        gp = grads_np(model.net, "pinn_")
        zero_grads(model)
        ld = I.DRM_loss(model, x)
        ld.backward()
        gd = grads_np(model.net, "drm_")
        save(tag, x=x.detach().numpy(), L=np.float64(L), n=np.int64(n), pinn_loss=np.float64(lp.item()),
             drm_loss=np.float64(ld.item()), **params_np(model.net, ""), **gp, **gd)

    torch.manual_seed(310)
    um = W.FCN([1, 20, 20, 1], L=L, enforce_bc=True).double()
    vm = W.FCN([1, 10, 10, 1], L=L, enforce_bc=False).double()
    with torch.no_grad():
        for m in linears(um.net) + linears(vm.net):
            m.bias.uniform_(-0.3, 0.3)
    n = 2
    x = torch.linspace(0.0, L, 65, dtype=torch.float64).view(-1, 1).requires_grad_(True)
    out = {}
    zero_grads(um, vm)
    total, lv, lpde, lnorm = W.WAN_loss(um, vm, x, n, L, 1.0, 1.0)
    total.backward(retain_graph=True)
    out.update({"tot_" + k: v for k, v in grads_np(um.net, "u_").items()})
    out.update({"tot_" + k: v for k, v in grads_np(vm.net, "v_").items()})
    zero_grads(um, vm)
    lv.backward()
    out.update({"lv_" + k: v for k, v in grads_np(um.net, "u_").items()})
    out.update({"lv_" + k: v for k, v in grads_np(vm.net, "v_").items()})
    save("ipw1d_wan_n2", x=x.detach().numpy(), L=np.float64(L), n=np.int64(n), total=np.float64(total.item()),
         loss_v=np.float64(lv.item()), loss_pde=np.float64(lpde.item()), loss_norm=np.float64(lnorm.item()),
         **params_np(um.net, "u_"), **params_np(vm.net, "v_"), **out)


# ------------------------------------------------------------------ 2-D oscillator
def qho2d_cases(Q):
    L = 6.0
    for tag, tech, nx, ny in (("qho2d_fbc_00", "FBC", 0, 0), ("qho2d_fn_21", "FN", 2, 1)):
        torch.manual_seed(400 + nx)
        model = Q.FCN([2, 16, 16, 16, 1], nx, ny, tech).double()
        g1 = torch.linspace(-L, L, 12, dtype=torch.float64)
        xg, yg = torch.meshgrid(g1, g1, indexing="ij")
        x = xg.clone().requires_grad_(True)
        y = yg.clone().requires_grad_(True)
        E = Q.Exact_energy(nx, ny, L)
        # --- restatement of the inline residual block, QHO_2D.py:329-341,363-383 ---
        def jets():
            u = model(x, y)
            ux = torch.autograd.grad(u, x, torch.ones_like(u), create_graph=True)[0]
            uy = torch.autograd.grad(u, y, torch.ones_like(u), create_graph=True)[0]
            return u, ux, uy
        zero_grads(model)
        u, ux, uy = jets()
        V = 0.5 * math.sqrt(2) ** 2 * (x ** 2 + y ** 2)
        uxx = torch.autograd.grad(ux, x, torch.ones_like(ux), create_graph=True)[0]
        uyy = torch.autograd.grad(uy, y, torch.ones_like(uy), create_graph=True)[0]
        res = -0.5 * (uxx + uyy) + V * u - E * u
        lp = torch.mean(res ** 2)
        lp.backward()
        gp = grads_np(model.net, "pinn_")
        zero_grads(model)
        u, ux, uy = jets()
        V = 0.5 * math.sqrt(2) ** 2 * (x ** 2 + y ** 2)
        ld = torch.mean(0.5 * (ux ** 2 + uy ** 2) + V * u ** 2) / torch.mean(u ** 2 + 1e-8)
        ld.backward()
        gd = grads_np(model.net, "drm_")
        save(tag, x=x.detach().numpy(), y=y.detach().numpy(), L=np.float64(L), nx=np.int64(nx), ny=np.int64(ny),
             E=np.float64(E), nodes_x=model.nodes_x.double().numpy(), nodes_y=model.nodes_y.double().numpy(),
             pinn_loss=np.float64(lp.item()), drm_loss=np.float64(ld.item()),
             u=u.detach().numpy(), **params_np(model.net, ""), **gp, **gd)

    torch.manual_seed(410)
    nx, ny = 1, 0
    um = Q.FCN([2, 16, 16, 1], nx, ny, "FBC").double()
    vm = Q.FCN([2, 10, 10, 1], nx, ny, "FBC").double()
    g1 = torch.linspace(-L, L, 12, dtype=torch.float64)
    xg, yg = torch.meshgrid(g1, g1, indexing="ij")
    x = xg.clone().requires_grad_(True)
    y = yg.clone().requires_grad_(True)
    out = {}
    zero_grads(um, vm)
    total, lv, lpde, lnorm = Q.WAN_loss(um, vm, x, y, nx, ny, L, 1.0, 1.0)
    total.backward(retain_graph=True)
    out.update({"tot_" + k: v for k, v in grads_np(um.net, "u_").items()})
    out.update({"tot_" + k: v for k, v in grads_np(vm.net, "v_").items()})
    zero_grads(um, vm)
    lv.backward()
    out.update({"lv_" + k: v for k, v in grads_np(um.net, "u_").items()})
    out.update({"lv_" + k: v for k, v in grads_np(vm.net, "v_").items()})
    save("qho2d_wan_10", x=x.detach().numpy(), y=y.detach().numpy(), L=np.float64(L), nx=np.int64(nx), ny=np.int64(ny),
         E=np.float64(Q.Exact_energy(nx, ny, L)), total=np.float64(total.item()), loss_v=np.float64(lv.item()),
         loss_pde=np.float64(lpde.item()), loss_norm=np.float64(lnorm.item()),
         **params_np(um.net, "u_"), **params_np(vm.net, "v_"), **out)


# ------------------------------------------------------------------ Kramers–Henneberger
def kh_cases(K):
    L, alpha, V0 = 12.0, 2.0, -24.856
    for tag, tech in (("kh1d_raw", "RAW"), ("kh1d_fbc", "FBC")):
        torch.manual_seed(500)
        model = K.UnifiedEigenModel([1, 16, 16, 1], technique=tech, E_init=-3.0, device="cpu").double()
        vm = K.FCN1D([1, 10, 10, 1], technique="RAW").double()
        x = torch.linspace(-L, L, 96, dtype=torch.float64).requires_grad_(True)
        Vx = K.V_KH(x.detach(), alpha=alpha, V0=V0, use_avg=True, n_theta=500)
        out = {}
        zero_grads(model, vm)
        lp = K.pinn_loss(model, x, alpha, V0)
        lp.backward()
        out.update(grads_np(model.u_model.net, "pinn_")); out["pinn_gE"] = model.energy.grad.numpy().copy()
        zero_grads(model, vm)
        ld = K.drm_loss(model, x, alpha, V0, L)
        ld.backward()
        out.update(grads_np(model.u_model.net, "drm_"))
        zero_grads(model, vm)
        lw, ln = K.wan_loss(model, vm, x, alpha, V0, L)
        (lw + ln).backward()
        out.update(grads_np(model.u_model.net, "wan_u_")); out.update(grads_np(vm.net, "wan_v_"))
        out["wan_gE"] = model.energy.grad.numpy().copy()
        save(tag, x=x.detach().numpy(), V=Vx.numpy(), L=np.float64(L), E=np.float64(model.energy.item()),
             pinn_loss=np.float64(lp.item()), drm_loss=np.float64(ld.item()),
             wan_pde=np.float64(lw.item()), wan_norm=np.float64(ln.item()),
             **params_np(model.u_model.net, "u_"), **params_np(vm.net, "v_"), **out)


# ------------------------------------------------------------------ remaining Schrödinger scripts
def perturb_biases(*seqs):
    with torch.no_grad():
        for seq in seqs:
            for m in seq:
                if isinstance(m, torch.nn.Linear):
                    m.bias.uniform_(-0.3, 0.3)


def qho1d_cases(Q1, QW):
    """QHO_1D_PINN_DRM.py (ModuleList sine network, exp-window / forced nodes) and QHO_1D_WAN.py
    (tanh networks, trainable energies)."""
    X_max = 6.0
    for tag, n, kw in (("qho1d_bc_n1", 1, dict(enforce_bc=True)), ("qho1d_fn_n2", 2, dict(enforce_bc=True, FN=True)),
                       ("qho1d_fnonly_n3", 3, dict(enforce_bc=False, FN=True))):
        torch.manual_seed(600 + n)
        model = Q1.FCN_Single([1, 20, 20, 1], num_states=n, domain_length=2 * X_max, **kw).double()
        x = torch.linspace(-X_max, X_max, 81, dtype=torch.float64).view(-1, 1).requires_grad_(True)
        out = {}
        for nm, fn in (("pinn", lambda: Q1.PINN_loss(model, x)), ("drm", lambda: Q1.DRM_loss(model, x)),
                       ("norm", lambda: Q1.normalization_loss(model, x)),
                       ("orth", lambda: Q1.Orthogonal_loss(model, x, n, X_max))):
            zero_grads(model)
            l = fn()
            l.backward()
            out[nm + "_loss"] = np.float64(l.item())
            out.update(grads_np(model.net.layers, nm + "_"))
        save(tag, x=x.detach().numpy(), X_max=np.float64(X_max), n=np.int64(n), **params_np(model.net.layers, ""), **out)

    torch.manual_seed(620)
    L, n = 6.0, 1
    um = QW.FCN([1, 20, 20, 1], num_states=n, L=L, enforce_bc=True).double()
    vm = QW.FCN([1, 10, 10, 1], num_states=n, L=L, enforce_bc=False).double()
    perturb_biases(um.net, vm.net)
    with torch.no_grad():
        um.energies.add_(0.1)
    x = torch.linspace(-L, L, 81, dtype=torch.float64).view(-1, 1).requires_grad_(True)
    out = {}
    zero_grads(um, vm)
    total, lv, lpde, lnorm = QW.WAN_loss(um, vm, x, n, L, 1.0, 1.0)
    total.backward(retain_graph=True)
    out.update({"tot_" + k: v for k, v in grads_np(um.net, "u_").items()})
    out.update({"tot_" + k: v for k, v in grads_np(vm.net, "v_").items()})
    out["tot_gE"] = um.energies.grad.numpy().copy()
    zero_grads(um, vm)
    lv.backward()
    out.update({"lv_" + k: v for k, v in grads_np(um.net, "u_").items()})
    out.update({"lv_" + k: v for k, v in grads_np(vm.net, "v_").items()})
    out["lv_gE"] = um.energies.grad.numpy().copy()
    save("qho1d_wan_n1", x=x.detach().numpy(), L=np.float64(L), n=np.int64(n), E=np.float64(um.energies.item()),
         total=np.float64(total.item()), loss_v=np.float64(lv.item()), loss_pde=np.float64(lpde.item()),
         loss_norm=np.float64(lnorm.item()), **params_np(um.net, "u_"), **params_np(vm.net, "v_"), **out)


def ipw_fn_wan_case(WF):
    """IPW_1D_WAN_FN.py: forced-node ansatz on both networks."""
    torch.manual_seed(630)
    L, n = 2.0, 3
    um = WF.FCN([1, 20, 20, 1], num_states=n, L=L).double()
    vm = WF.FCN([1, 10, 10, 1], num_states=1, L=L).double()
    perturb_biases(um.net, vm.net)
    x = torch.linspace(0.0, L, 65, dtype=torch.float64).view(-1, 1).requires_grad_(True)
    out = {}
    zero_grads(um, vm)
    total, lv, lpde, lnorm = WF.WAN_loss(um, vm, x, n, L, 1.0, 1.0)
    total.backward(retain_graph=True)
    out.update({"tot_" + k: v for k, v in grads_np(um.net, "u_").items()})
    out.update({"tot_" + k: v for k, v in grads_np(vm.net, "v_").items()})
    zero_grads(um, vm)
    lv.backward()
    out.update({"lv_" + k: v for k, v in grads_np(um.net, "u_").items()})
    out.update({"lv_" + k: v for k, v in grads_np(vm.net, "v_").items()})
    save("ipw1d_wanfn_n3", x=x.detach().numpy(), L=np.float64(L), n=np.int64(n), total=np.float64(total.item()),
         loss_v=np.float64(lv.item()), loss_pde=np.float64(lpde.item()), loss_norm=np.float64(lnorm.item()),
         **params_np(um.net, "u_"), **params_np(vm.net, "v_"), **out)


def ipw2d_cases(I2):
    """IPW_2D.py: the inline PINN / DRM blocks of train_pinn_seperate (:195-228) restated around the
    imported FCN, plus orthogonal_loss (:113-126)."""
    L = 2.0
    for tag, tech, nx, ny in (("ipw2d_fbc_11", "FBC", 1, 1), ("ipw2d_fn_32", "FN", 3, 2)):
        torch.manual_seed(640 + nx)
        model = I2.FCN([2, 16, 16, 16, 1], nx, ny, tech).double()
        g1 = torch.linspace(0.0, L, 12, dtype=torch.float64)
        xg, yg = torch.meshgrid(g1, g1, indexing="ij")
        x = xg.clone().requires_grad_(True)
        y = yg.clone().requires_grad_(True)
        k2 = (2 * ((nx * np.pi) ** 2 / (2 * L ** 2) + (ny * np.pi) ** 2 / (2 * L ** 2)))

        def jets():
            u = model(x, y)
            ux = torch.autograd.grad(u, x, torch.ones_like(u), create_graph=True)[0]
            uy = torch.autograd.grad(u, y, torch.ones_like(u), create_graph=True)[0]
            return u, ux, uy
        zero_grads(model)
        u, ux, uy = jets()
        uxx = torch.autograd.grad(ux, x, torch.ones_like(ux), create_graph=True)[0]
        uyy = torch.autograd.grad(uy, y, torch.ones_like(uy), create_graph=True)[0]
        lp = torch.mean((uxx + uyy + k2 * u) ** 2)
        lp.backward()
        gp = grads_np(model.net, "pinn_")
        zero_grads(model)
        u, ux, uy = jets()
        ld = torch.mean(ux ** 2 + uy ** 2) / torch.mean(u ** 2 + 1e-8)
        ld.backward()
        gd = grads_np(model.net, "drm_")
        zero_grads(model)
        lo = I2.orthogonal_loss(model, x, y, nx, ny, L)
        go = {}
        if torch.is_tensor(lo) and lo.requires_grad:
            lo.backward()
            go = grads_np(model.net, "orth_")
        save(tag, x=x.detach().numpy(), y=y.detach().numpy(), L=np.float64(L), nx=np.int64(nx), ny=np.int64(ny),
             pinn_loss=np.float64(lp.item()), drm_loss=np.float64(ld.item()), orth_loss=np.float64(float(lo)),
             u=u.detach().numpy(), **params_np(model.net, ""), **gp, **gd, **go)


def qho2d_energy_case(QE):
    """QHO_2D_Energy.py: PINN residual with the trainable energy E_train (:287-291,382-383)."""
    L, nx, ny = 6.0, 1, 1
    torch.manual_seed(650)
    model = QE.FCN([2, 16, 16, 16, 1], nx, ny, "FBC").double()
    E_train = torch.nn.Parameter(torch.tensor(QE.Exact_energy(nx, ny, L) + 0.05, dtype=torch.float64))
    g1 = torch.linspace(-L, L, 12, dtype=torch.float64)
    xg, yg = torch.meshgrid(g1, g1, indexing="ij")
    x = xg.clone().requires_grad_(True)
    y = yg.clone().requires_grad_(True)
    u = model(x, y)
    ux = torch.autograd.grad(u, x, torch.ones_like(u), create_graph=True)[0]
    uy = torch.autograd.grad(u, y, torch.ones_like(u), create_graph=True)[0]
    uxx = torch.autograd.grad(ux, x, torch.ones_like(ux), create_graph=True)[0]
    uyy = torch.autograd.grad(uy, y, torch.ones_like(uy), create_graph=True)[0]
    V = 0.5 * math.sqrt(2) ** 2 * (x ** 2 + y ** 2)
    lp = torch.mean((-0.5 * (uxx + uyy) + V * u - E_train * u) ** 2)
    zero_grads(model)
    lp.backward()
    save("qho2d_energy_11", x=x.detach().numpy(), y=y.detach().numpy(), L=np.float64(L), nx=np.int64(nx), ny=np.int64(ny),
         E=np.float64(E_train.item()), pinn_loss=np.float64(lp.item()), pinn_gE=E_train.grad.numpy().copy(),
         **params_np(model.net, ""), **grads_np(model.net, "pinn_"))


def main():
    torch.set_default_dtype(torch.float32)
    P = load_ref("Poisson_Equations/Poisson_ND.py", "ref_poisson_nd")
    poisson_cases(P)
    I = load_ref("Schrodinger_Equations/Infinite_Potential_Well/IPW_1D_PINN_DRM.py", "ref_ipw_pd")
    W = load_ref("Schrodinger_Equations/Infinite_Potential_Well/IPW_1D_WAN.py", "ref_ipw_wan")
    ipw_cases(I, W)
    Q = load_ref("Schrodinger_Equations/Quantum_Harmonic_Oscillator/QHO_2D.py", "ref_qho2d")
    qho2d_cases(Q)
    K = load_ref("Schrodinger_Equations/Kramers_Henneberger/KH_1D.py", "ref_kh1d")
    kh_cases(K)
    more()


def more():
    """The scripts added after the first fixture set (run alone with ``make_golden.py more``)."""
    torch.set_default_dtype(torch.float32)
    Q1 = load_ref("Schrodinger_Equations/Quantum_Harmonic_Oscillator/QHO_1D_PINN_DRM.py", "ref_qho1d_pd")
    QW = load_ref("Schrodinger_Equations/Quantum_Harmonic_Oscillator/QHO_1D_WAN.py", "ref_qho1d_wan")
    qho1d_cases(Q1, QW)
    WF = load_ref("Schrodinger_Equations/Infinite_Potential_Well/IPW_1D_WAN_FN.py", "ref_ipw_wanfn")
    ipw_fn_wan_case(WF)
    I2 = load_ref("Schrodinger_Equations/Infinite_Potential_Well/IPW_2D.py", "ref_ipw2d")
    ipw2d_cases(I2)
    QE = load_ref("Schrodinger_Equations/Quantum_Harmonic_Oscillator/QHO_2D_Energy.py", "ref_qho2d_energy")
    qho2d_energy_case(QE)


if __name__ == "__main__":
    more() if sys.argv[1:] == ["more"] else main()
