"""GPU parity tests of the Schrödinger drop-ins (pde_b200.schrodinger.*) against the golden
fixtures produced by the live reference scripts (tests/golden/make_golden.py): every script, on small networks
and at the shapes BASELINE.json configs 4 and 5 name (cfg4_* / cfg5_* fixtures: 200 x 200 grids with
[2,50,50,50,50,1], 1000 / 1024-point grids with [1,50,50,50,1] / [1,100,100,100,1] / [1,200,200,200,1]).

Bars (conftest.loss_bar / grads_bar): 1e-10 relative in float64, always; 1e-5 in float32, raised only to twice
the reference's own float32-vs-float64 distance on the same fixture where that is larger."""
import numpy as np
import pytest
import torch

from conftest import assert_grads_golden, assert_loss_close, load_golden, net_from

pytestmark = pytest.mark.gpu
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


def _layers(Ws):
    return [Ws[0].shape[1]] + [W.shape[0] for W in Ws]


def _grid(g, dtype):
    """(x, y) mesh of a 2-D fixture: stored explicitly, or rebuilt from the stored size (config-shaped fixtures
    hold `grid_n` only; make_golden.grid2d: linspace incl. end points, indexing 'ij')."""
    if "x" in g:
        x, y = torch.tensor(g["x"]), torch.tensor(g["y"])
    else:
        L = float(g["L"])
        lo = -L if "E" in g else 0.0      # oscillator grids are [-L, L]^2, well grids [0, L]^2
        g1 = torch.linspace(lo, L, int(g["grid_n"]), dtype=torch.float64)
        x, y = torch.meshgrid(g1, g1, indexing="ij")
    mk = lambda t: t.to("cuda", dtype).clone().requires_grad_(True)
    return mk(x), mk(y)


def _check_wan(g, dtype, outs, ul, vl, energy=None, backward_second=True):
    """(total, loss_v, loss_pde, loss_norm) of a two-network WAN loss, gradients of `total` and of `loss_v`."""
    total, lv, lpde, lnorm = outs
    for got, key in ((total, "total"), (lv, "loss_v"), (lpde, "loss_pde"), (lnorm, "loss_norm")):
        assert_loss_close(got, g, key, dtype)
    total.backward(retain_graph=True)
    assert_grads_golden(_grads(ul), g, "tot_u_", dtype)
    assert_grads_golden(_grads(vl), g, "tot_v_", dtype)
    if energy is not None:
        assert_loss_close(energy.grad, g, "tot_gE", dtype)
        energy.grad = None
    for l in ul + vl:
        l.weight.grad = None; l.bias.grad = None
    lv.backward()
    assert_grads_golden(_grads(ul), g, "lv_u_", dtype)
    assert_grads_golden(_grads(vl), g, "lv_v_", dtype)
    if energy is not None:
        assert_loss_close(energy.grad, g, "lv_gE", dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name", ["ipw1d_fbc_n2", "ipw1d_fn_n3"])
def test_ipw1d_pinn_drm(name, dtype):
    from pde_b200.schrodinger import ipw_1d_pinn_drm as I
    g = load_golden(name)
    Ws, bs = net_from(g)
    L, n = float(g["L"]), int(g["n"])
    kw = dict(FN=True, num_states=3) if name.endswith("fn_n3") else dict(enforce_bc=True)
    model = I.FCN(_layers(Ws), L=L, **kw).double()
    lin = _load(model.net, Ws, bs)
    model = model.to("cuda", dtype)
    x = torch.tensor(g["x"], dtype=dtype, device="cuda", requires_grad=True)
    lp = I.PINN_loss(model, x, n, L); lp.backward()
    assert_loss_close(lp, g, "pinn_loss", dtype, name)
    assert_grads_golden(_grads(lin), g, "pinn_", dtype, name)
    _zero(model)
    ld = I.DRM_loss(model, x); ld.backward()
    assert_loss_close(ld, g, "drm_loss", dtype, name)
    assert_grads_golden(_grads(lin), g, "drm_", dtype, name)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name", ["ipw1d_wan_n2", "cfg5_ipw1d_wan_n2"])
def test_ipw1d_wan(name, dtype):
    """IPW_1D_WAN.py; cfg5_*: u [1,50,50,50,1] / v [1,20,20,20,1] on linspace(0, 2, 1000) (BASELINE config 5)."""
    from pde_b200.schrodinger import ipw_1d_wan as W
    g = load_golden(name)
    L, n = float(g["L"]), int(g["n"])
    um = W.FCN(_layers(net_from(g, "u_")[0]), L=L, enforce_bc=True).double()
    vm = W.FCN(_layers(net_from(g, "v_")[0]), L=L, enforce_bc=False).double()
    ul = _load(um.net, *net_from(g, "u_")); vl = _load(vm.net, *net_from(g, "v_"))
    um, vm = um.to("cuda", dtype), vm.to("cuda", dtype)
    x = torch.tensor(g["x"], dtype=dtype, device="cuda", requires_grad=True)
    _check_wan(g, dtype, W.WAN_loss(um, vm, x, n, L, 1.0, 1.0), ul, vl)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name,tech", [("qho2d_fbc_00", "FBC"), ("qho2d_fn_21", "FN"),
                                       ("cfg4_qho2d_fbc_00", "FBC"), ("cfg4_qho2d_fn_21", "FN")])
def test_qho2d_pinn_drm(name, tech, dtype):
    """QHO_2D.py inline PINN / Rayleigh blocks; cfg4_*: [2,50,50,50,50,1] on the 200 x 200 grid (BASELINE config 4)."""
    from pde_b200.schrodinger import qho_2d as Q
    g = load_golden(name)
    L, nx, ny, E = float(g["L"]), int(g["nx"]), int(g["ny"]), float(g["E"])
    Ws, bs = net_from(g)
    model = Q.FCN(_layers(Ws), nx, ny, tech).double()
    np.testing.assert_allclose(model.nodes_x.double().numpy(), g["nodes_x"], rtol=0, atol=0)
    lin = _load(model.net, Ws, bs)
    model = model.to("cuda", dtype)
    x, y = _grid(g, dtype)
    lp = Q.PINN_loss(model, x, y, E, L); lp.backward()
    assert_loss_close(lp, g, "pinn_loss", dtype, name)
    assert_grads_golden(_grads(lin), g, "pinn_", dtype, name)
    _zero(model)
    ld = Q.DRM_loss(model, x, y, L); ld.backward()
    assert_loss_close(ld, g, "drm_loss", dtype, name)
    assert_grads_golden(_grads(lin), g, "drm_", dtype, name)
    with pytest.raises(ValueError):
        model.technique = "XX"
        Q.PINN_loss(model, x, y, E, L)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name", ["qho2d_wan_10", "cfg4_qho2d_wan_10"])
def test_qho2d_wan(name, dtype):
    """QHO_2D.WAN_loss incl. its finite-norm term (4 L^2 mean u^2 - 1)^2 (:222); cfg4_*: u [2,50,50,50,50,1] /
    v [2,20,20,20,1] on the 200 x 200 grid."""
    from pde_b200.schrodinger import qho_2d as Q
    g = load_golden(name)
    L, nx, ny = float(g["L"]), int(g["nx"]), int(g["ny"])
    um = Q.FCN(_layers(net_from(g, "u_")[0]), nx, ny, "FBC").double()
    vm = Q.FCN(_layers(net_from(g, "v_")[0]), nx, ny, "FBC").double()
    ul = _load(um.net, *net_from(g, "u_")); vl = _load(vm.net, *net_from(g, "v_"))
    um, vm = um.to("cuda", dtype), vm.to("cuda", dtype)
    x, y = _grid(g, dtype)
    _check_wan(g, dtype, Q.WAN_loss(um, vm, x, y, nx, ny, L, 1.0, 1.0), ul, vl)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name,tech", [("kh1d_raw", "RAW"), ("kh1d_fbc", "FBC"), ("cfg5_kh1d_a10", "FBC")])
def test_kh1d(name, tech, dtype):
    """KH_1D.py pinn / drm / wan losses with the trainable energy; cfg5_*: u [1,100,100,100,1] ('FBC' for PINN / DRM,
    'RAW' for WAN as train_state_v2 builds it, :331) / v [1,50,50,50,1] on linspace(-60, 60, 1024), alpha = 10."""
    from pde_b200.schrodinger import kh_1d as K
    g = load_golden(name)
    L, V0 = float(g["L"]), -24.856
    alpha = float(g["alpha"]) if "alpha" in g else 2.0
    model = K.UnifiedEigenModel(_layers(net_from(g, "u_")[0]), technique=tech, E_init=float(g["E"])).double()
    vm = K.FCN1D(_layers(net_from(g, "v_")[0]), technique="RAW").double()
    ul = _load(model.u_model.net, *net_from(g, "u_")); vl = _load(vm.net, *net_from(g, "v_"))
    wmodel, wl = model, ul
    if "w_W0" in g:    # separate RAW network for the WAN loss
        wmodel = K.UnifiedEigenModel(_layers(net_from(g, "w_")[0]), technique="RAW", E_init=float(g["E"])).double()
        wl = _load(wmodel.u_model.net, *net_from(g, "w_"))
        wmodel = wmodel.to("cuda", dtype)
    model, vm = model.to("cuda", dtype), vm.to("cuda", dtype)
    x = torch.tensor(g["x"], dtype=dtype, device="cuda", requires_grad=True)
    Vx = K.V_KH(x.detach(), alpha=alpha, V0=V0)
    assert np.max(np.abs(Vx.double().cpu().numpy() - g["V"])) <= (1e-11 if dtype == torch.float64 else 2e-5)
    lp = K.pinn_loss(model, x, alpha, V0); lp.backward()
    assert_loss_close(lp, g, "pinn_loss", dtype, name)
    assert_grads_golden(_grads(ul), g, "pinn_", dtype, name)
    assert_loss_close(model.energy.grad, g, "pinn_gE", dtype, name)
    _zero(model, vm)
    ld = K.drm_loss(model, x, alpha, V0, L); ld.backward()
    assert_loss_close(ld, g, "drm_loss", dtype, name)
    assert_grads_golden(_grads(ul), g, "drm_", dtype, name)
    _zero(model, vm, wmodel)
    lw, ln = K.wan_loss(wmodel, vm, x, alpha, V0, L)
    assert_loss_close(lw, g, "wan_pde", dtype, name); assert_loss_close(ln, g, "wan_norm", dtype, name)
    (lw + ln).backward()
    assert_grads_golden(_grads(wl), g, "wan_u_", dtype, name)
    assert_grads_golden(_grads(vl), g, "wan_v_", dtype, name)
    assert_loss_close(wmodel.energy.grad, g, "wan_gE", dtype, name)


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
    for nm, fn in (("pinn", lambda: Q.PINN_loss(model, x)), ("drm", lambda: Q.DRM_loss(model, x)),
                   ("norm", lambda: Q.normalization_loss(model, x)), ("orth", lambda: Q.Orthogonal_loss(model, x, n, X_max))):
        _zero(model)
        l = fn(); l.backward()
        assert_loss_close(l, g, nm + "_loss", dtype, name)
        assert_grads_golden(_grads(lin), g, nm + "_", dtype, name)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name", ["qho1d_wan_n1", "cfg5_qho1d_wan_n1"])
def test_qho1d_wan(name, dtype):
    """QHO_1D_WAN.py: trainable ``energies`` gets its gradient from the same launches; cfg5_*: u [1,200,200,200,1] /
    v [1,100,100,100,1] on linspace(-6, 6, 1000) (QHO_1D_WAN.py:159,169-176)."""
    from pde_b200.schrodinger import qho_1d_wan as W
    g = load_golden(name)
    L, n = float(g["L"]), int(g["n"])
    cfg = name.startswith("cfg5")
    um = W.FCN(_layers(net_from(g, "u_")[0]), num_states=n, L=L, enforce_bc=True).double()
    vm = W.FCN(_layers(net_from(g, "v_")[0]), num_states=n, L=L, enforce_bc=cfg).double()
    ul = _load(um.net, *net_from(g, "u_")); vl = _load(vm.net, *net_from(g, "v_"))
    with torch.no_grad():
        um.energies.fill_(float(g["E"]))
    um, vm = um.to("cuda", dtype), vm.to("cuda", dtype)
    x = torch.tensor(g["x"], dtype=dtype, device="cuda", requires_grad=True)
    _check_wan(g, dtype, W.WAN_loss(um, vm, x, n, L, 1.0, 1.0), ul, vl, energy=um.energies)


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
    _check_wan(g, dtype, W.WAN_loss(um, vm, x, n, L, 1.0, 1.0), ul, vl)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name,tech", [("ipw2d_fbc_11", "FBC"), ("ipw2d_fn_32", "FN"),
                                       ("cfg4_ipw2d_fbc_11", "FBC"), ("cfg4_ipw2d_fn_32", "FN")])
def test_ipw2d(name, tech, dtype):
    """IPW_2D.py inline PINN / Rayleigh blocks and orthogonal_loss; cfg4_*: [2,50,50,50,50,1], 200 x 200 grid."""
    from pde_b200.schrodinger import ipw_2d as I
    g = load_golden(name)
    L, nx, ny = float(g["L"]), int(g["nx"]), int(g["ny"])
    Ws, bs = net_from(g)
    model = I.FCN(_layers(Ws), nx, ny, tech).double()
    lin = _load(model.net, Ws, bs)
    model = model.to("cuda", dtype)
    x, y = _grid(g, dtype)
    lp = I.PINN_loss(model, x, y, nx, ny, L); lp.backward()
    assert_loss_close(lp, g, "pinn_loss", dtype, name)
    assert_grads_golden(_grads(lin), g, "pinn_", dtype, name)
    _zero(model)
    ld = I.DRM_loss(model, x, y, L); ld.backward()
    assert_loss_close(ld, g, "drm_loss", dtype, name)
    assert_grads_golden(_grads(lin), g, "drm_", dtype, name)
    _zero(model)
    lo = I.orthogonal_loss(model, x, y, nx, ny, L)
    assert_loss_close(lo, g, "orth_loss", dtype, name)
    if "orth_gW0" in g:
        lo.backward()
        assert_grads_golden(_grads(lin), g, "orth_", dtype, name)
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
    x, y = _grid(g, dtype)
    lp = Q.PINN_loss(model, x, y, E_train, L); lp.backward()
    assert_loss_close(lp, g, "pinn_loss", dtype)
    assert_grads_golden(_grads(lin), g, "pinn_", dtype, "qho2d energy")
    assert_loss_close(E_train.grad, g, "pinn_gE", dtype)


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
