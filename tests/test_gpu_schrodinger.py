"""GPU parity tests of the Schrödinger drop-ins (pde_b200.schrodinger.*) against the golden
fixtures produced by the live reference scripts (tests/golden/make_golden.py): IPW 1-D PINN / DRM
(hard-BC and forced-node ansatz), IPW 1-D WAN, QHO 2-D PINN / DRM / WAN, KH 1-D PINN / DRM / WAN with
trainable energy.  fp64 at 1e-10-level, fp32 at 1e-5-level (BASELINE.json north_star); the WAN and
quotient losses, whose gradients are differences of gradient vectors, get 4x those bars as in
tests/test_gpu_poisson.py."""
import numpy as np
import pytest
import torch

from conftest import assert_grads_close, grads_from, load_golden, net_from

pytestmark = pytest.mark.gpu
TOL = {torch.float64: 2e-10, torch.float32: 1e-5}
DTYPES = [torch.float64, torch.float32]


def _load(net, Ws, bs):
    lin = [m for m in net if isinstance(m, torch.nn.Linear)]
    with torch.no_grad():
        for l, W, b in zip(lin, Ws, bs):
            l.weight.copy_(torch.tensor(W)); l.bias.copy_(torch.tensor(b))
    return lin


def _grads(lin):
    z = lambda p: (p.grad if p.grad is not None else torch.zeros_like(p)).double().cpu().numpy()
    return [z(l.weight) for l in lin], [z(l.bias) for l in lin]


def _zero(*mods):
    for m in mods:
        for p in m.parameters():
            p.grad = None


def _close(a, b, tol):
    a = float(a.detach()) if torch.is_tensor(a) else float(a)
    assert abs(a - float(b)) <= tol * max(abs(float(b)), 1e-3), (a, float(b))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name", ["ipw1d_fbc_n2", "ipw1d_fn_n3"])
def test_ipw1d_pinn_drm(name, dtype):
    from pde_b200.schrodinger import ipw_1d_pinn_drm as I
    g = load_golden(name)
    Ws, bs = net_from(g)
    L, n = float(g["L"]), int(g["n"])
    kw = dict(FN=True, num_states=3) if name.endswith("fn_n3") else dict(enforce_bc=True)
    model = I.FCN([1, 20, 20, 1], L=L, **kw).double()
    lin = _load(model.net, Ws, bs)
    model = model.to("cuda", dtype)
    x = torch.tensor(g["x"], dtype=dtype, device="cuda", requires_grad=True)
    tol = TOL[dtype]
    lp = I.PINN_loss(model, x, n, L); lp.backward()
    _close(lp, g["pinn_loss"], tol)
    assert_grads_close(_grads(lin), grads_from(g, "pinn_"), tol, name + " pinn")
    _zero(model)
    ld = I.DRM_loss(model, x); ld.backward()
    _close(ld, g["drm_loss"], tol)
    assert_grads_close(_grads(lin), grads_from(g, "drm_"), 4 * tol, name + " drm")


@pytest.mark.parametrize("dtype", DTYPES)
def test_ipw1d_wan(dtype):
    from pde_b200.schrodinger import ipw_1d_wan as W
    g = load_golden("ipw1d_wan_n2")
    L, n = float(g["L"]), int(g["n"])
    um = W.FCN([1, 20, 20, 1], L=L, enforce_bc=True).double()
    vm = W.FCN([1, 10, 10, 1], L=L, enforce_bc=False).double()
    ul = _load(um.net, *net_from(g, "u_")); vl = _load(vm.net, *net_from(g, "v_"))
    um, vm = um.to("cuda", dtype), vm.to("cuda", dtype)
    x = torch.tensor(g["x"], dtype=dtype, device="cuda", requires_grad=True)
    tol = 4 * TOL[dtype]
    total, lv, lpde, lnorm = W.WAN_loss(um, vm, x, n, L, 1.0, 1.0)
    for got, key in ((total, "total"), (lv, "loss_v"), (lpde, "loss_pde"), (lnorm, "loss_norm")):
        _close(got, g[key], tol)
    total.backward(retain_graph=True)
    assert_grads_close(_grads(ul), grads_from(g, "tot_u_"), tol, "tot/u")
    assert_grads_close(_grads(vl), grads_from(g, "tot_v_"), tol, "tot/v")
    _zero(um, vm)
    lv.backward()
    assert_grads_close(_grads(ul), grads_from(g, "lv_u_"), tol, "lv/u")
    assert_grads_close(_grads(vl), grads_from(g, "lv_v_"), tol, "lv/v")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name,tech", [("qho2d_fbc_00", "FBC"), ("qho2d_fn_21", "FN")])
def test_qho2d_pinn_drm(name, tech, dtype):
    from pde_b200.schrodinger import qho_2d as Q
    g = load_golden(name)
    L, nx, ny, E = float(g["L"]), int(g["nx"]), int(g["ny"]), float(g["E"])
    model = Q.FCN([2, 16, 16, 16, 1], nx, ny, tech).double()
    np.testing.assert_allclose(model.nodes_x.double().numpy(), g["nodes_x"], rtol=0, atol=0)
    lin = _load(model.net, *net_from(g))
    model = model.to("cuda", dtype)
    x = torch.tensor(g["x"], dtype=dtype, device="cuda", requires_grad=True)
    y = torch.tensor(g["y"], dtype=dtype, device="cuda", requires_grad=True)
    tol = TOL[dtype]
    lp = Q.PINN_loss(model, x, y, E, L); lp.backward()
    _close(lp, g["pinn_loss"], tol)
    assert_grads_close(_grads(lin), grads_from(g, "pinn_"), tol, name + " pinn")
    _zero(model)
    ld = Q.DRM_loss(model, x, y, L); ld.backward()
    _close(ld, g["drm_loss"], tol)
    assert_grads_close(_grads(lin), grads_from(g, "drm_"), 4 * tol, name + " drm")
    with pytest.raises(ValueError):
        model.technique = "XX"
        Q.PINN_loss(model, x, y, E, L)


@pytest.mark.parametrize("dtype", DTYPES)
def test_qho2d_wan(dtype):
    from pde_b200.schrodinger import qho_2d as Q
    g = load_golden("qho2d_wan_10")
    L, nx, ny = float(g["L"]), int(g["nx"]), int(g["ny"])
    um = Q.FCN([2, 16, 16, 1], nx, ny, "FBC").double()
    vm = Q.FCN([2, 10, 10, 1], nx, ny, "FBC").double()
    ul = _load(um.net, *net_from(g, "u_")); vl = _load(vm.net, *net_from(g, "v_"))
    um, vm = um.to("cuda", dtype), vm.to("cuda", dtype)
    x = torch.tensor(g["x"], dtype=dtype, device="cuda", requires_grad=True)
    y = torch.tensor(g["y"], dtype=dtype, device="cuda", requires_grad=True)
    tol = 4 * TOL[dtype]
    total, lv, lpde, lnorm = Q.WAN_loss(um, vm, x, y, nx, ny, L, 1.0, 1.0)
    for got, key in ((total, "total"), (lv, "loss_v"), (lpde, "loss_pde"), (lnorm, "loss_norm")):
        _close(got, g[key], 10 * tol)     # (4 L^2 mean u^2 - 1)^2 amplifies the rounding of mean u^2
    total.backward(retain_graph=True)
    assert_grads_close(_grads(ul), grads_from(g, "tot_u_"), 10 * tol, "tot/u")
    assert_grads_close(_grads(vl), grads_from(g, "tot_v_"), 10 * tol, "tot/v")
    _zero(um, vm)
    lv.backward()
    assert_grads_close(_grads(ul), grads_from(g, "lv_u_"), 10 * tol, "lv/u")
    assert_grads_close(_grads(vl), grads_from(g, "lv_v_"), 10 * tol, "lv/v")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name,tech", [("kh1d_raw", "RAW"), ("kh1d_fbc", "FBC")])
def test_kh1d(name, tech, dtype):
    from pde_b200.schrodinger import kh_1d as K
    g = load_golden(name)
    L, alpha, V0 = float(g["L"]), 2.0, -24.856
    model = K.UnifiedEigenModel([1, 16, 16, 1], technique=tech, E_init=float(g["E"])).double()
    vm = K.FCN1D([1, 10, 10, 1], technique="RAW").double()
    ul = _load(model.u_model.net, *net_from(g, "u_")); vl = _load(vm.net, *net_from(g, "v_"))
    model, vm = model.to("cuda", dtype), vm.to("cuda", dtype)
    x = torch.tensor(g["x"], dtype=dtype, device="cuda", requires_grad=True)
    Vx = K.V_KH(x.detach(), alpha=alpha, V0=V0)
    assert np.max(np.abs(Vx.double().cpu().numpy() - g["V"])) <= (1e-11 if dtype == torch.float64 else 2e-5)
    tol = TOL[dtype] if dtype == torch.float64 else 4 * TOL[dtype]   # fp32: V itself is rounded (n_theta mean)
    lp = K.pinn_loss(model, x, alpha, V0); lp.backward()
    _close(lp, g["pinn_loss"], tol)
    assert_grads_close(_grads(ul), grads_from(g, "pinn_"), tol, name + " pinn")
    _close(model.energy.grad, g["pinn_gE"], tol)
    _zero(model, vm)
    ld = K.drm_loss(model, x, alpha, V0, L); ld.backward()
    _close(ld, g["drm_loss"], tol)
    assert_grads_close(_grads(ul), grads_from(g, "drm_"), 4 * tol, name + " drm")
    _zero(model, vm)
    lw, ln = K.wan_loss(model, vm, x, alpha, V0, L)
    _close(lw, g["wan_pde"], 10 * tol); _close(ln, g["wan_norm"], 10 * tol)
    (lw + ln).backward()
    assert_grads_close(_grads(ul), grads_from(g, "wan_u_"), 10 * tol, name + " wan/u")
    assert_grads_close(_grads(vl), grads_from(g, "wan_v_"), 10 * tol, name + " wan/v")
    _close(model.energy.grad, g["wan_gE"], 10 * tol)


# ---------------------------------------------------------------- second fixture set: the remaining scripts
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name,kw", [("qho1d_bc_n1", dict(enforce_bc=True)), ("qho1d_fn_n2", dict(enforce_bc=True, FN=True)),
                                     ("qho1d_fnonly_n3", dict(enforce_bc=False, FN=True))])
def test_qho1d_pinn_drm(name, kw, dtype):
    """QHO_1D_PINN_DRM.py: ModuleList sine network; PINN, DRM, normalisation and orthogonality terms."""
    from pde_b200.schrodinger import qho_1d_pinn_drm as Q
    g = load_golden(name)
    X_max, n = float(g["X_max"]), int(g["n"])
    model = Q.FCN_Single([1, 20, 20, 1], num_states=n, domain_length=2 * X_max, **kw).double()
    lin = _load(model.net.layers, *net_from(g))
    model = model.to("cuda", dtype)
    x = torch.tensor(g["x"], dtype=dtype, device="cuda", requires_grad=True)
    tol = TOL[dtype]
    for nm, fn, k in (("pinn", lambda: Q.PINN_loss(model, x), 1), ("drm", lambda: Q.DRM_loss(model, x), 4),
                      ("norm", lambda: Q.normalization_loss(model, x), 4),
                      ("orth", lambda: Q.Orthogonal_loss(model, x, n, X_max), 4)):
        _zero(model)
        l = fn(); l.backward()
        _close(l, g[nm + "_loss"], k * tol)
        assert_grads_close(_grads(lin), grads_from(g, nm + "_"), k * tol, f"{name} {nm}")


@pytest.mark.parametrize("dtype", DTYPES)
def test_qho1d_wan(dtype):
    """QHO_1D_WAN.py: trainable ``energies`` gets its gradient from the same launches."""
    from pde_b200.schrodinger import qho_1d_wan as W
    g = load_golden("qho1d_wan_n1")
    L, n = float(g["L"]), int(g["n"])
    um = W.FCN([1, 20, 20, 1], num_states=n, L=L, enforce_bc=True).double()
    vm = W.FCN([1, 10, 10, 1], num_states=n, L=L, enforce_bc=False).double()
    ul = _load(um.net, *net_from(g, "u_")); vl = _load(vm.net, *net_from(g, "v_"))
    with torch.no_grad():
        um.energies.fill_(float(g["E"]))
    um, vm = um.to("cuda", dtype), vm.to("cuda", dtype)
    x = torch.tensor(g["x"], dtype=dtype, device="cuda", requires_grad=True)
    tol = 10 * 4 * TOL[dtype]
    total, lv, lpde, lnorm = W.WAN_loss(um, vm, x, n, L, 1.0, 1.0)
    for got, key in ((total, "total"), (lv, "loss_v"), (lpde, "loss_pde"), (lnorm, "loss_norm")):
        _close(got, g[key], tol)
    total.backward(retain_graph=True)
    assert_grads_close(_grads(ul), grads_from(g, "tot_u_"), tol, "tot/u")
    assert_grads_close(_grads(vl), grads_from(g, "tot_v_"), tol, "tot/v")
    _close(um.energies.grad, g["tot_gE"], tol)
    _zero(um, vm)
    lv.backward()
    assert_grads_close(_grads(ul), grads_from(g, "lv_u_"), tol, "lv/u")
    assert_grads_close(_grads(vl), grads_from(g, "lv_v_"), tol, "lv/v")
    _close(um.energies.grad, g["lv_gE"], tol)


@pytest.mark.parametrize("dtype", DTYPES)
def test_ipw1d_wan_fn(dtype):
    from pde_b200.schrodinger import ipw_1d_wan_fn as W
    g = load_golden("ipw1d_wanfn_n3")
    L, n = float(g["L"]), int(g["n"])
    um = W.FCN([1, 20, 20, 1], num_states=n, L=L).double()
    vm = W.FCN([1, 10, 10, 1], num_states=1, L=L).double()
    ul = _load(um.net, *net_from(g, "u_")); vl = _load(vm.net, *net_from(g, "v_"))
    um, vm = um.to("cuda", dtype), vm.to("cuda", dtype)
    x = torch.tensor(g["x"], dtype=dtype, device="cuda", requires_grad=True)
    tol = 4 * TOL[dtype]
    total, lv, lpde, lnorm = W.WAN_loss(um, vm, x, n, L, 1.0, 1.0)
    for got, key in ((total, "total"), (lv, "loss_v"), (lpde, "loss_pde"), (lnorm, "loss_norm")):
        _close(got, g[key], tol)
    total.backward(retain_graph=True)
    assert_grads_close(_grads(ul), grads_from(g, "tot_u_"), tol, "tot/u")
    assert_grads_close(_grads(vl), grads_from(g, "tot_v_"), tol, "tot/v")
    _zero(um, vm)
    lv.backward()
    assert_grads_close(_grads(ul), grads_from(g, "lv_u_"), tol, "lv/u")
    assert_grads_close(_grads(vl), grads_from(g, "lv_v_"), tol, "lv/v")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name,tech", [("ipw2d_fbc_11", "FBC"), ("ipw2d_fn_32", "FN")])
def test_ipw2d(name, tech, dtype):
    from pde_b200.schrodinger import ipw_2d as I
    g = load_golden(name)
    L, nx, ny = float(g["L"]), int(g["nx"]), int(g["ny"])
    model = I.FCN([2, 16, 16, 16, 1], nx, ny, tech).double()
    lin = _load(model.net, *net_from(g))
    model = model.to("cuda", dtype)
    x = torch.tensor(g["x"], dtype=dtype, device="cuda", requires_grad=True)
    y = torch.tensor(g["y"], dtype=dtype, device="cuda", requires_grad=True)
    tol = TOL[dtype]
    lp = I.PINN_loss(model, x, y, nx, ny, L); lp.backward()
    _close(lp, g["pinn_loss"], tol)
    assert_grads_close(_grads(lin), grads_from(g, "pinn_"), tol, name + " pinn")
    _zero(model)
    ld = I.DRM_loss(model, x, y, L); ld.backward()
    _close(ld, g["drm_loss"], tol)
    assert_grads_close(_grads(lin), grads_from(g, "drm_"), 4 * tol, name + " drm")
    _zero(model)
    lo = I.orthogonal_loss(model, x, y, nx, ny, L)
    _close(lo, g["orth_loss"], 4 * tol)
    if "orth_gW0" in g:
        lo.backward()
        assert_grads_close(_grads(lin), grads_from(g, "orth_"), 4 * tol, name + " orth")
    with pytest.raises(ValueError):
        model.technique = "XX"
        I.PINN_loss(model, x, y, nx, ny, L)


@pytest.mark.parametrize("dtype", DTYPES)
def test_qho2d_trainable_energy(dtype):
    """QHO_2D_Energy.py: E_train.grad from the fused PINN launch."""
    from pde_b200.schrodinger import qho_2d_energy as Q
    g = load_golden("qho2d_energy_11")
    L, nx, ny = float(g["L"]), int(g["nx"]), int(g["ny"])
    model = Q.FCN([2, 16, 16, 16, 1], nx, ny, "FBC").double()
    lin = _load(model.net, *net_from(g))
    model = model.to("cuda", dtype)
    E_train = torch.nn.Parameter(torch.tensor(float(g["E"]), dtype=dtype, device="cuda"))
    x = torch.tensor(g["x"], dtype=dtype, device="cuda", requires_grad=True)
    y = torch.tensor(g["y"], dtype=dtype, device="cuda", requires_grad=True)
    tol = TOL[dtype]
    lp = Q.PINN_loss(model, x, y, E_train, L); lp.backward()
    _close(lp, g["pinn_loss"], tol)
    assert_grads_close(_grads(lin), grads_from(g, "pinn_"), tol, "qho2d energy pinn")
    _close(E_train.grad, g["pinn_gE"], tol)


@pytest.mark.parametrize("method", ["PINN", "DRM"])
def test_ipw2d_lbfgs_closure(method):
    """IPW_2D.py:169-170,271-312 polishes with torch.optim.LBFGS(strong_wolfe) through a closure that
    re-evaluates the whole loss.  The drop-in losses are ordinary autograd ops, so the same closure works on
    them: 12 L-BFGS iterations (many closure calls each) from identical weights end at the same parameters
    as the reference closure (nested autograd on the CPU, restated inline like the reference does), fp64."""
    from pde_b200.schrodinger import ipw_2d as I
    L, nx, ny = 2.0, 1, 1
    torch.manual_seed(3)
    ref = I.FCN([2, 16, 16, 1], nx, ny, "FBC").double()
    ours = I.FCN([2, 16, 16, 1], nx, ny, "FBC").double()
    ours.load_state_dict(ref.state_dict())
    ours = ours.cuda()
    g = torch.linspace(0.0, L, 24, dtype=torch.float64)
    xs, ys = torch.meshgrid(g, g, indexing="ij")
    Xd, Yd = torch.rand(64, dtype=torch.float64) * L, torch.rand(64, dtype=torch.float64) * L
    ud = I.Exact_solution(L, nx, ny, Xd, Yd)
    k2 = I.k_squared(nx, ny, L)

    def ref_closure_factory(opt):
        x = xs.clone().requires_grad_(True); y = ys.clone().requires_grad_(True)

        def closure():
            opt.zero_grad()
            u = ref(x, y, L)
            ux = torch.autograd.grad(u, x, torch.ones_like(u), create_graph=True)[0]
            uy = torch.autograd.grad(u, y, torch.ones_like(u), create_graph=True)[0]
            if method == "PINN":
                uxx = torch.autograd.grad(ux, x, torch.ones_like(ux), create_graph=True)[0]
                uyy = torch.autograd.grad(uy, y, torch.ones_like(uy), create_graph=True)[0]
                main = torch.mean((uxx + uyy + k2 * u) ** 2)
            else:
                main = torch.mean(ux ** 2 + uy ** 2) / torch.mean(u ** 2 + 1e-8)
            total = main + 10.0 * torch.mean((ref(Xd, Yd, L) - ud) ** 2)
            total.backward()
            return total
        return closure

    def our_closure_factory(opt):
        x, y = xs.cuda(), ys.cuda()
        Xc, Yc, uc = Xd.cuda(), Yd.cuda(), ud.cuda()

        def closure():
            opt.zero_grad()
            main = I.PINN_loss(ours, x, y, nx, ny, L) if method == "PINN" else I.DRM_loss(ours, x, y, L)
            total = main + 10.0 * I.data_loss(ours, Xc, Yc, uc, L)
            total.backward()
            return total
        return closure

    finals = []
    for model, factory in ((ref, ref_closure_factory), (ours, our_closure_factory)):
        opt = torch.optim.LBFGS(model.parameters(), lr=0.5, max_iter=12, line_search_fn="strong_wolfe")
        loss0 = opt.step(factory(opt))
        finals.append((float(loss0.detach()), [p.detach().cpu().clone() for p in model.parameters()]))
    assert abs(finals[0][0] - finals[1][0]) <= 1e-9 * max(abs(finals[0][0]), 1e-3)
    for a, b in zip(finals[0][1], finals[1][1]):
        assert float((a - b).abs().max()) <= 1e-6 * max(1.0, float(a.abs().max())), "L-BFGS trajectories diverged"
