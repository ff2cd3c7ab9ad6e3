"""GPU parity tests (run with -m gpu on a B200): CUDA path through the C ABI vs the golden
fixtures produced by the live reference and vs the numpy oracle on seeded inputs.

Tolerances (BASELINE.json north_star): 1e-5 relative in fp32, 1e-10 in fp64, on losses and on
parameter gradients (per-tensor relative L2 and global max-abs)."""
import math

import numpy as np
import pytest
import torch

from conftest import (assert_grads_close, assert_grads_golden, assert_loss_close, grads_from, load_golden, net_from)

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 1e-5, torch.float64: 1e-10}

POISSON = [
    ("poisson_pinn_d1_w64_fbc", "pinn", "FBC"), ("poisson_pinn_d3_w64_fbc", "pinn", "FBC"),
    ("poisson_drm_d5_w64_rb", "drm", "RB"), ("poisson_pinn_d2_w16_fbc", "pinn", "FBC"),
    ("poisson_pinn_d5_w16_fbc", "pinn", "FBC"), ("poisson_pinn_d3_w16_rb", "pinn", "RB"),
    ("poisson_pinn_d4_w12_fbc", "pinn", "FBC"), ("poisson_drm_d1_w16_fbc", "drm", "FBC"),
    ("poisson_drm_d2_w16_fbc", "drm", "FBC"), ("poisson_drm_d3_w16_fbc", "drm", "FBC"),
]


def _model_from(pb, g, bc, dtype, prefix="", cls=None):
    Ws, bs = net_from(g, prefix)
    d, w, depth = Ws[0].shape[1], Ws[0].shape[0], len(Ws)
    m = (cls or pb.poisson.SolutionNet)(d, w, depth, bc) if cls is None else cls(d, w, depth)
    m = m.double()
    lin = [x for x in m.net if isinstance(x, torch.nn.Linear)]
    with torch.no_grad():
        for l, W, b in zip(lin, Ws, bs):
            l.weight.copy_(torch.tensor(W)); l.bias.copy_(torch.tensor(b))
    return m.to("cuda", dtype)


def _grads_of(m):
    lin = [x for x in m.net if isinstance(x, torch.nn.Linear)]
    gW = [(l.weight.grad if l.weight.grad is not None else torch.zeros_like(l.weight)).double().cpu().numpy() for l in lin]
    gb = [(l.bias.grad if l.bias.grad is not None else torch.zeros_like(l.bias)).double().cpu().numpy() for l in lin]
    return gW, gb


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("name,method,bc", POISSON)
def test_poisson_loss_and_grads_vs_reference_golden(name, method, bc, dtype):
    import pde_b200 as pb
    g = load_golden(name)
    m = _model_from(pb, g, bc, dtype)
    X = torch.tensor(g["X"], dtype=dtype, device="cuda", requires_grad=True)
    f = torch.tensor(g["f"], dtype=dtype, device="cuda")
    fn = pb.poisson.pinn_residual_loss if method == "pinn" else pb.poisson.drm_energy_loss
    loss = fn(m, X, f, float(g["L"]))
    loss.backward()
    tol = TOL[dtype]
    assert abs(loss.item() - g["loss"]) <= tol * max(abs(g["loss"]), 1e-3), (loss.item(), g["loss"])
    assert_grads_close(_grads_of(m), grads_from(g), tol, name)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("name,method,bc", POISSON[:7])
def test_jets_vs_reference_golden(name, method, bc, dtype):
    """u, grad u, Laplacian from the jet kernel equal the reference helpers' outputs."""
    import pde_b200 as pb
    g = load_golden(name)
    m = _model_from(pb, g, bc, dtype)
    X = torch.tensor(g["X"], dtype=dtype, device="cuda")
    u, gr, h = pb.poisson.solution_jets(m, X, float(g["L"]), 2)
    tol = TOL[dtype] * 5  # pointwise values, no averaging
    for got, want in ((u, g["u"]), (gr, g["grad_u"]), (h.sum(1, keepdim=True), g["lap_u"])):
        want = np.asarray(want)
        assert np.max(np.abs(got.detach().double().cpu().numpy() - want)) <= tol * max(1.0, np.abs(want).max())


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_jets_backward_matches_oracle(dtype):
    """Reverse sweep with arbitrary cotangents vs oracle/jets_numpy.mlp_jets_backward."""
    import pde_b200 as pb
    from oracle import jets_numpy as O
    rng = np.random.default_rng(5)
    for d, w, depth, order, act, N in [(3, 64, 5, 2, "sin", 333), (2, 50, 4, 2, "tanh", 257), (1, 20, 3, 1, "tanh", 100),
                                       (5, 24, 3, 1, "sin", 64), (2, 200, 4, 0, "sin", 40), (4, 9, 2, 2, "sin", 33)]:
        Ws = [rng.uniform(-1, 1, (w, d)) / math.sqrt(d)] + [rng.uniform(-1, 1, (w, w)) / math.sqrt(w) for _ in range(depth - 2)] \
            + [rng.uniform(-1, 1, (1, w)) / math.sqrt(w)]
        bs = [rng.uniform(-0.5, 0.5, W.shape[0]) for W in Ws]
        X = rng.uniform(0, 2, (N, d)); Jbar = rng.normal(size=(N, 1 + order * d))
        A = O.SIN if act == "sin" else O.TANH
        J, cache = O.mlp_jets_forward(Ws, bs, X, A, order)
        gWs, gbs = O.mlp_jets_backward(Ws, bs, cache, Jbar)
        mods = []
        for i in range(depth - 1):
            mods += [torch.nn.Linear(Ws[i].shape[1], w), pb.poisson.Sin() if act == "sin" else torch.nn.Tanh()]
        mods += [torch.nn.Linear(w, 1)]
        net = torch.nn.Sequential(*mods).double()
        lin = [x for x in net if isinstance(x, torch.nn.Linear)]
        with torch.no_grad():
            for l, W, b in zip(lin, Ws, bs):
                l.weight.copy_(torch.tensor(W)); l.bias.copy_(torch.tensor(b))
        net = net.to("cuda", dtype)
        Jg = pb.mlp_jets(net, torch.tensor(X, dtype=dtype, device="cuda"), order)
        tol = TOL[dtype]
        assert np.max(np.abs(Jg.detach().double().cpu().numpy() - J)) <= 5 * tol * max(1.0, np.abs(J).max())
        Jg.backward(torch.tensor(Jbar, dtype=dtype, device="cuda"))
        gW = [l.weight.grad.double().cpu().numpy() for l in lin]
        gb = [l.bias.grad.double().cpu().numpy() for l in lin]
        assert_grads_close((gW, gb), (gWs, gbs), tol, f"d{d} w{w} {act} order{order}")


def test_poisson_wan_vs_reference_golden():
    import pde_b200 as pb
    g = load_golden("poisson_wan_d2_w16")
    for dtype in (torch.float64, torch.float32):
        um = _model_from(pb, g, "FBC", dtype, "u_")
        vm = _model_from(pb, g, None, dtype, "v_", cls=pb.poisson.CriticNet)
        X = torch.tensor(g["X"], dtype=dtype, device="cuda", requires_grad=True)
        f = torch.tensor(g["f"], dtype=dtype, device="cuda")
        lu, lv, weak, pn = pb.poisson.wan_losses(um, vm, X, f, float(g["L"]), v_reg_weight=float(g["v_reg_weight"]))
        for got, key in ((lu, "loss_u"), (lv, "loss_v"), (weak, "weak"), (pn, "phi_norm")):
            assert_loss_close(got, g, key, dtype, "poisson wan")
        lu.backward(retain_graph=True)
        assert_grads_golden(_grads_of(um), g, "lu_u_", dtype)
        assert_grads_golden(_grads_of(vm), g, "lu_v_", dtype)
        um.zero_grad(); vm.zero_grad()
        lv.backward()
        assert_grads_golden(_grads_of(um), g, "lv_u_", dtype)
        assert_grads_golden(_grads_of(vm), g, "lv_v_", dtype)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_value_terms_vs_reference_golden(dtype):
    """The value-only terms of the reference epoch (Poisson_ND.py:130-147,230-239) on the points the fixture
    stores: data MSE (data_loss), both norm_loss modes on u from the order-0 jet kernel, and the Dirichlet face
    penalty face by face (what boundary_loss_dirichlet evaluates on its own draws)."""
    import pde_b200 as pb
    g = load_golden("poisson_value_terms_d3_w16_rb")
    L = float(g["L"])
    m = _model_from(pb, g, "RB", dtype)
    Xd = torch.tensor(g["Xd"], dtype=dtype, device="cuda")
    ud = torch.tensor(g["ud"], dtype=dtype, device="cuda")
    ld = pb.poisson.data_loss(m, Xd, ud, L); ld.backward()
    assert_loss_close(ld, g, "data_loss", dtype)
    assert_grads_golden(_grads_of(m), g, "data_", dtype)
    for mode in ("nontrivial", "l2"):
        m.zero_grad()
        u = pb.poisson.solution_jets(m, Xd, L, 0)[0]
        ln = pb.poisson.norm_loss(u, mode=mode); ln.backward()
        assert_loss_close(ln, g, f"norm_{mode}_loss", dtype)
        assert_grads_golden(_grads_of(m), g, f"norm_{mode}_", dtype)
    with pytest.raises(ValueError):
        pb.poisson.norm_loss(u, mode="xx")
    m.zero_grad()
    Xb = torch.tensor(g["Xb"], dtype=dtype, device="cuda")
    lb = sum(pb.poisson.data_loss(m, xb, None, L) for xb in Xb) / Xb.shape[0]
    lb.backward()
    assert_loss_close(lb, g, "bc_loss", dtype)
    assert_grads_golden(_grads_of(m), g, "bc_", dtype)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("dim,bc", [(2, "RB"), (3, "FBC"), (5, "RB")])
def test_boundary_loss_dirichlet_vs_oracle(dim, bc, dtype):
    """boundary_loss_dirichlet draws its 2d face batches from the device generator (Poisson_ND.py:130-141: the
    global RNG there).  Re-seeding replays the draws, and the numpy oracle evaluates mean over faces of mean(u^2)
    and its parameter gradient on exactly those points."""
    import pde_b200 as pb
    from oracle import jets_numpy as O
    L, Nb = 2.0, 37
    torch.manual_seed(11 + dim)
    m = pb.poisson.SolutionNet(dim, 24, 4, bc).to("cuda", dtype)
    torch.manual_seed(5)
    lb = pb.poisson.boundary_loss_dirichlet(m, L, Nb, dim, "cuda")
    lb.backward()
    torch.manual_seed(5)
    faces = []
    for i in range(dim):
        for at_L in (False, True):
            X = torch.rand(Nb, dim, device="cuda", dtype=dtype) * L
            X[:, i] = L if at_L else 0.0
            faces.append(X.double().cpu().numpy())
    lin = [x for x in m.net if isinstance(x, torch.nn.Linear)]
    Ws = [l.weight.detach().double().cpu().numpy() for l in lin]
    bs = [l.bias.detach().double().cpu().numpy() for l in lin]
    env = dict(kind=O.ENV_POLY if bc == "FBC" else O.ENV_NONE, lo=0.0, hi=L)
    want, gW, gb = 0.0, [np.zeros_like(W) for W in Ws], [np.zeros_like(b) for b in bs]
    for X in faces:
        l, gWs, gbs = O.mse_loss(Ws, bs, X, O.SIN, env, None)
        want += l / len(faces)
        for a, b in zip(gW, gWs):
            a += b / len(faces)
        for a, b in zip(gb, gbs):
            a += b / len(faces)
    tol = TOL[dtype]
    if bc == "FBC":   # the envelope vanishes on every face: the term and its gradient are exactly zero
        assert abs(lb.item()) <= 1e-12 and want == 0.0
    else:
        assert abs(lb.item() - want) <= tol * abs(want), (lb.item(), want)
        assert_grads_close(_grads_of(m), (gW, gb), tol, f"bc d{dim}")


def test_chunk_linearity_and_ragged_sizes():
    """Size-independent property at a non-multiple-of-tile N: loss/grad of the whole batch equal the
    N-weighted combination of two uneven chunks (plain-mean losses)."""
    import pde_b200 as pb
    torch.manual_seed(0)
    for dtype, tol in ((torch.float64, 1e-11), (torch.float32, 2e-5)):
        m = pb.poisson.SolutionNet(3, 64, 5, "FBC").to("cuda", dtype)
        N, cut = 4099, 1237
        X = torch.rand(N, 3, device="cuda", dtype=dtype) * 2
        f = pb.poisson.rhs_f_for_u_sin(X, 2.0, [1, 1, 1])

        def run(Xc, fc):
            m.zero_grad()
            l = pb.poisson.pinn_residual_loss(m, Xc, fc, 2.0)
            l.backward()
            return l.item(), torch.cat([p.grad.reshape(-1) for p in m.parameters()]).double()
        l_all, g_all = run(X, f)
        l_a, g_a = run(X[:cut], f[:cut])
        l_b, g_b = run(X[cut:], f[cut:])
        l_mix = (cut * l_a + (N - cut) * l_b) / N
        g_mix = (cut * g_a + (N - cut) * g_b) / N
        assert abs(l_all - l_mix) <= tol * abs(l_all)
        assert (g_all - g_mix).abs().max().item() <= tol * g_all.abs().max().item()


def test_single_point_and_no_grad():
    import pde_b200 as pb
    m = pb.poisson.SolutionNet(2, 16, 3, "FBC").cuda()
    X = torch.rand(1, 2, device="cuda") * 2
    f = torch.ones(1, 1, device="cuda")
    with torch.no_grad():
        l = pb.poisson.pinn_residual_loss(m, X, f, 2.0)
    assert math.isfinite(l.item()) and not l.requires_grad
    with pytest.raises(ValueError):
        m.bc_mode = "XX"
        pb.poisson.pinn_residual_loss(m, X, f, 2.0)
    with pytest.raises(pb.PdeError):
        pb.poisson.pinn_residual_loss(pb.poisson.SolutionNet(2, 16, 3), X.cpu(), f.cpu(), 2.0)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("kind", [0, 1])
def test_wan_scalar_losses_vs_reference_formulas(kind, dtype):
    """pde_wan_scalars against the tensor expressions the reference's WAN losses end with (Poisson_ND.py:118-127,
    IPW_1D_WAN.py:108-114; kind 1: KH_1D.py:263-268), values and gradients with respect to the four means."""
    from pde_b200.ops import wan_scalar_losses
    torch.manual_seed(kind)
    for trial in range(4):
        m = (torch.rand(4, dtype=torch.float64) + 0.05) * torch.tensor([1.0, 0.3, 0.02, 2.0], dtype=torch.float64)
        m[0] = m[0] * (-1) ** trial
        vol, reg, wp, wn, e1, e2 = 4.0 + trial, 0.7, 10.0, 3.0, (1e-8 if kind == 0 else 1e-12), 1e-8
        mr = m.clone().requires_grad_(True)
        if kind == 0:
            pde = mr[0] ** 2 / (mr[1] + e1)
        else:
            pde = (vol * mr[0] / (vol * mr[1] + e1)) ** 2
        lv = -torch.log(pde + e2) + reg * mr[3]
        nrm = (vol * mr[2] - 1.0) ** 2
        tot = wp * pde + wn * nrm
        mg = m.to("cuda", dtype).requires_grad_(True)
        o = wan_scalar_losses(mg, kind=kind, eps_pde=e1, eps_log=e2, vol=vol, reg=reg, w_pde=wp, w_norm=wn)
        tol = 1e-12 if dtype == torch.float64 else 2e-6
        for got, want in zip(o, (pde, lv, nrm, tot)):
            assert abs(float(got.detach()) - float(want.detach())) <= tol * max(1.0, abs(float(want.detach())))
        for got, want in ((o[3], tot), (o[1], lv), (o[0] + 2.0 * o[2], pde + 2.0 * nrm)):
            gw, = torch.autograd.grad(want, mr, retain_graph=True)
            gg, = torch.autograd.grad(got, mg, retain_graph=True)
            assert float((gg.double().cpu() - gw).abs().max()) <= tol * max(1.0, float(gw.abs().max()))
