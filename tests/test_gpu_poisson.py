"""GPU parity tests (run with -m gpu on a B200): CUDA path through the C ABI vs the golden
fixtures produced by the live reference and vs the numpy oracle on seeded inputs.

Tolerances (BASELINE.json north_star): 1e-5 relative in fp32, 1e-10 in fp64, on losses and on
parameter gradients (per-tensor relative L2 and global max-abs)."""
import math

import numpy as np
import pytest
import torch

from conftest import assert_grads_close, grads_from, load_golden, net_from

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
        tol = TOL[dtype]
        lu, lv, weak, pn = pb.poisson.wan_losses(um, vm, X, f, float(g["L"]), v_reg_weight=float(g["v_reg_weight"]))
        for got, key in ((lu, "loss_u"), (lv, "loss_v"), (weak, "weak"), (pn, "phi_norm")):
            assert abs(got.item() - g[key]) <= 4 * tol * max(abs(g[key]), 1e-3), key
        lu.backward(retain_graph=True)
        assert_grads_close(_grads_of(um), grads_from(g, "lu_u_"), 4 * tol, "lu/u")
        assert_grads_close(_grads_of(vm), grads_from(g, "lu_v_"), 4 * tol, "lu/v")
        um.zero_grad(); vm.zero_grad()
        lv.backward()
        assert_grads_close(_grads_of(um), grads_from(g, "lv_u_"), 4 * tol, "lv/u")
        assert_grads_close(_grads_of(vm), grads_from(g, "lv_v_"), 4 * tol, "lv/v")


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
