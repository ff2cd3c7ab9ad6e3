"""GPU parity tests of the tcgen05 path (run with -m gpu on a B200).

The tensor-core kernel is forced with PDE_B200_PATH=tc (it is the default above 4096 points) and
checked, through the C ABI, against (i) the golden fixtures produced by the live reference,
(ii) the numpy oracle on seeded inputs covering activations, dimensions, programs, envelopes,
padding and multi-tile cases, and (iii) size-independent properties at large N.
Tolerance: 1e-5 relative in fp32 (BASELINE.json north_star) on losses and parameter gradients."""
import math
import os

import numpy as np
import pytest
import torch

from conftest import assert_grads_close, grads_from, load_golden, net_from

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(autouse=True)
def _force_tc():
    import pde_b200 as pb
    with pb.ops.kernel_path("tc"):
        yield


def _seq(Ws, bs, act, dtype=torch.float32):
    import pde_b200 as pb
    mods = []
    for i in range(len(Ws) - 1):
        mods += [torch.nn.Linear(Ws[i].shape[1], Ws[i].shape[0]), pb.poisson.Sin() if act == "sin" else torch.nn.Tanh()]
    mods += [torch.nn.Linear(Ws[-1].shape[1], 1)]
    net = torch.nn.Sequential(*mods).double()
    lin = [x for x in net if isinstance(x, torch.nn.Linear)]
    with torch.no_grad():
        for l, W, b in zip(lin, Ws, bs):
            l.weight.copy_(torch.tensor(W)); l.bias.copy_(torch.tensor(b))
    return net.to("cuda", dtype), lin


def _rand_net(rng, d, w, depth):
    Ws = [rng.uniform(-1, 1, (w, d)) / math.sqrt(d)] + [rng.uniform(-1, 1, (w, w)) / math.sqrt(w) for _ in range(depth - 2)] \
        + [rng.uniform(-1, 1, (1, w)) / math.sqrt(w)]
    bs = [rng.uniform(-0.5, 0.5, W.shape[0]) for W in Ws]
    f32 = lambda a: a.astype(np.float32).astype(np.float64)   # the kernel sees fp32 parameters
    return [f32(W) for W in Ws], [f32(b) for b in bs]


def _grads(lin):
    return [l.weight.grad.double().cpu().numpy() for l in lin], [l.bias.grad.double().cpu().numpy() for l in lin]


GOLDEN = [("poisson_pinn_d5_w16_fbc", "pinn", "FBC"),      # 7 jet channels: the dimension-split path (two passes + pointwise kernel)
          ("poisson_pinn_d1_w64_fbc", "pinn", "FBC"), ("poisson_pinn_d3_w64_fbc", "pinn", "FBC"),
          ("poisson_drm_d5_w64_rb", "drm", "RB"), ("poisson_pinn_d2_w16_fbc", "pinn", "FBC"),
          ("poisson_pinn_d3_w16_rb", "pinn", "RB"), ("poisson_pinn_d4_w12_fbc", "pinn", "FBC"),
          ("poisson_drm_d1_w16_fbc", "drm", "FBC"), ("poisson_drm_d2_w16_fbc", "drm", "FBC"),
          ("poisson_drm_d3_w16_fbc", "drm", "FBC")]


@pytest.mark.parametrize("name,method,bc", GOLDEN)
def test_tc_vs_reference_golden(name, method, bc):
    import pde_b200 as pb
    from pde_b200 import _lib as L
    g = load_golden(name)
    Ws, bs = net_from(g)
    m = pb.poisson.SolutionNet(Ws[0].shape[1], Ws[0].shape[0], len(Ws), bc).double()
    lin = [x for x in m.net if isinstance(x, torch.nn.Linear)]
    with torch.no_grad():
        for l, W, b in zip(lin, Ws, bs):
            l.weight.copy_(torch.tensor(W)); l.bias.copy_(torch.tensor(b))
    m = m.to("cuda", torch.float32)
    X = torch.tensor(g["X"], dtype=torch.float32, device="cuda", requires_grad=True)
    f = torch.tensor(g["f"], dtype=torch.float32, device="cuda")
    fn = pb.poisson.pinn_residual_loss if method == "pinn" else pb.poisson.drm_energy_loss
    loss = fn(m, X, f, float(g["L"]))
    assert pb.ops.last_kernel_path() == "tcgen05"
    loss.backward()
    assert abs(loss.item() - g["loss"]) <= TOL * max(abs(g["loss"]), 1e-3), (loss.item(), g["loss"])
    assert_grads_close(_grads(lin), grads_from(g), TOL, name)


CASES = [  # d, width, depth, act, program, envelope, N
    (3, 64, 5, "sin", "pinn", "poly", 64), (3, 64, 5, "sin", "pinn", "poly", 1000),
    (3, 64, 5, "sin", "pinn", "poly", 64 * 148 * 2 + 17), (5, 64, 5, "sin", "drm", "none", 3000),
    (1, 64, 5, "sin", "pinn", "poly", 777), (2, 50, 5, "sin", "pinn", "exp", 2049),
    (1, 50, 4, "tanh", "pinn", "poly", 1000), (2, 50, 5, "sin", "rayleigh", "poly", 1500),
    (4, 33, 3, "tanh", "pinn", "none", 500), (2, 20, 4, "tanh", "drm", "poly", 129), (3, 64, 3, "sin", "mse", "poly", 4097),
    # five channels at every depth the kernel takes: 2 / 3 / 4 hidden layers = 1 / 2 / 3 GEMM layers behind the two W
    # slots and the chunk-0 shadow (depth 3: the first layer follows the top layer directly); three tiles per CTA
    (3, 64, 3, "tanh", "pinn", "poly", 900), (3, 40, 4, "sin", "pinn", "poly", 1300), (3, 64, 5, "sin", "pinn", "poly", 64 * 148 * 3 + 5),
]


def _reference_fp32_error_rayleigh(Ws, bs, X, beta, act, a, L_box, draws=4):
    """Largest gradient error (conftest.grads_err) of the reference's algorithm — nested torch autograd, oracle/autograd_ref
    helpers, x (L - x) envelope, mean(a |grad u|^2 + beta u^2) / mean(u^2) — run in float32 against its own float64 run."""
    from conftest import grads_err
    from oracle import autograd_ref as AR
    gen = np.random.default_rng(1)

    def run(Wl, bl, Xa, ba, dtype):
        net = AR.build_mlp([Wl[0].shape[1]] + [W.shape[0] for W in Wl], act, dtype)
        AR.load_params(net, Wl, bl)
        Xt = torch.tensor(Xa, dtype=dtype).requires_grad_(True)
        u = AR.solution(net, Xt, L_box, "FBC")
        g = AR._grad(u, Xt)
        bt = torch.tensor(ba, dtype=dtype)
        loss = (a * (g * g).sum(dim=1, keepdim=True) + bt * u * u).mean() / (u * u).mean()
        loss.backward()
        lin = [m for m in net if isinstance(m, torch.nn.Linear)]
        return [m.weight.grad.double().numpy() for m in lin], [m.bias.grad.double().numpy() for m in lin]
    worst = 0.0
    for k in range(draws + 1):
        jit = (lambda t: t * (1.0 + gen.uniform(-1, 1, t.shape) * 2.0 ** -24)) if k else (lambda t: t)
        Wl, bl, Xa, ba = [jit(W) for W in Ws], [jit(b) for b in bs], jit(X), jit(beta)
        worst = max(worst, grads_err(run(Wl, bl, Xa, ba, torch.float32), run(Wl, bl, Xa, ba, torch.float64))[0])
    return worst


@pytest.mark.parametrize("d,w,depth,act,prog,env_kind,N", CASES)
def test_tc_vs_numpy_oracle(d, w, depth, act, prog, env_kind, N):
    import pde_b200 as pb
    from pde_b200 import _lib as L
    from pde_b200.ops import EnvelopeSpec, ProgramSpec, residual_means
    from oracle import jets_numpy as O
    rng = np.random.default_rng(d * 1000 + w + N)
    Ws, bs = _rand_net(rng, d, w, depth)
    net, lin = _seq(Ws, bs, act)
    f32 = lambda a: a.astype(np.float32).astype(np.float64)
    X = f32(rng.uniform(0.05, 1.95, (N, d))); f = f32(rng.normal(size=(N, 1))); beta = f32(rng.uniform(0.5, 1.5, (N, 1)))
    A = O.SIN if act == "sin" else O.TANH
    env = {"kind": {"poly": O.ENV_POLY, "exp": O.ENV_EXPWIN, "none": O.ENV_NONE}[env_kind], "lo": 0.0, "hi": 2.0}
    espec = EnvelopeSpec({"poly": L.ENV_POLY, "exp": L.ENV_EXPWIN, "none": L.ENV_NONE}[env_kind], 0.0, 2.0)
    Xg, fg, bg = (torch.tensor(a, dtype=torch.float32, device="cuda") for a in (X, f, beta))
    if prog == "pinn":
        want, gWs, gbs, _ = O.eigen_pinn_loss(Ws, bs, X, A, env, -1.0, beta, 0.3, f=f)
        loss = residual_means(net, Xg, ProgramSpec(L.PROG_PINN, -1.0, 0.0, 0.3), espec, f=fg, beta=bg)[0]
    elif prog == "drm":
        def program(U):
            q, Ubar = O.drm_poisson_program(U, d, f)
            return float(q.mean()), Ubar / N, None
        want, gWs, gbs, _ = O._loss_and_grads(Ws, bs, X, A, 1, env, program)
        loss = residual_means(net, Xg, ProgramSpec(L.PROG_DRM, 0.5), espec, f=fg)[0]
    elif prog == "rayleigh":
        want, gWs, gbs, _ = O.rayleigh_loss(Ws, bs, X, A, env, 0.5, beta)
        m = residual_means(net, Xg, ProgramSpec(L.PROG_RAYLEIGH, 0.5), espec, beta=bg)
        loss = m[0] / m[1]
    else:
        def program(U):
            r = U[:, 0:1] - f
            Ubar = np.zeros_like(U); Ubar[:, 0:1] = 2 * r / N
            return float((r * r).mean()), Ubar, None
        want, gWs, gbs, _ = O._loss_and_grads(Ws, bs, X, A, 0, env, program)
        loss = residual_means(net, Xg, ProgramSpec(L.PROG_MSE), espec, f=fg)[0]
    assert pb.ops.last_kernel_path() == "tcgen05"
    loss.backward()
    # losses that are sums of cancelling terms (Deep Ritz) are compared on the scale of their terms
    assert abs(loss.item() - want) <= TOL * max(abs(want), 1e-2), (loss.item(), want)
    bar = TOL
    if prog == "rayleigh":
        # the quotient's gradient is a difference of nearly parallel vectors; the bar follows the rule of the golden
        # tests (conftest.grads_bar): max(1e-5, 2 x the float32 error of the reference's own nested-autograd algorithm
        # on this very case, worst of the case itself and four inputs one float32 ulp away)
        bar = max(TOL, 2.0 * _reference_fp32_error_rayleigh(Ws, bs, X, beta, act, 0.5, 2.0))
    assert_grads_close(_grads(lin), (gWs, gbs), bar, f"d{d} w{w} {act} {prog}")


def test_tc_large_batch_against_fp64_generic_kernel():
    """2^20 points: fp32 tensor-core result vs the generic kernel run in fp64 on the same points.
    Guards the accumulation over ~7000 tiles per CTA (the tensor core truncates when accumulating;
    running sums are kept in round-to-nearest fp32 outside it)."""
    import pde_b200 as pb
    from pde_b200 import _lib as L
    from pde_b200.ops import EnvelopeSpec, ProgramSpec, residual_means
    rng = np.random.default_rng(3)
    Ws, bs = _rand_net(rng, 3, 64, 5)
    n32, lin32 = _seq(Ws, bs, "sin")
    n64, lin64 = _seq(Ws, bs, "sin", torch.float64)
    N = 1 << 20
    X32 = torch.rand(N, 3, device="cuda") * 1.9 + 0.05
    f32 = torch.randn(N, device="cuda")
    espec = EnvelopeSpec(L.ENV_POLY, 0.0, 2.0)
    l32 = residual_means(n32, X32, ProgramSpec(L.PROG_PINN, -1.0), espec, f=f32)[0]
    assert pb.ops.last_kernel_path() == "tcgen05"
    l32.backward()
    with pb.ops.kernel_path("simt"):
        l64 = residual_means(n64, X32.double(), ProgramSpec(L.PROG_PINN, -1.0), espec, f=f32.double())[0]
        l64.backward()
    assert abs(l32.item() - l64.item()) <= TOL * abs(l64.item())
    assert_grads_close(_grads(lin32), _grads(lin64), TOL, "2^20 points")


def test_tc_chunk_linearity_full_size_and_determinism():
    """Size-independent properties at the benchmark's 2^22 points per GPU: the loss / gradient of the
    whole batch equal the N-weighted combination of two uneven chunks, and a repeated launch is
    bit-identical (fixed-order reductions, no atomics on the result path)."""
    import pde_b200 as pb
    torch.manual_seed(0)
    m = pb.poisson.SolutionNet(3, 64, 5, "FBC").cuda()
    N, cut = 1 << 22, (1 << 21) + 12345
    X = torch.rand(N, 3, device="cuda") * 2
    f = pb.poisson.rhs_f_for_u_sin(X, 2.0, [1, 1, 1])

    def run(Xc, fc):
        m.zero_grad()
        l = pb.poisson.pinn_residual_loss(m, Xc, fc, 2.0)
        l.backward()
        return l.item(), torch.cat([p.grad.reshape(-1) for p in m.parameters()]).double()
    l_all, g_all = run(X, f)
    assert pb.ops.last_kernel_path() == "tcgen05"
    l_again, g_again = run(X, f)
    assert l_all == l_again and torch.equal(g_all, g_again)
    l_a, g_a = run(X[:cut], f[:cut])
    l_b, g_b = run(X[cut:], f[cut:])
    l_mix = (cut * l_a + (N - cut) * l_b) / N
    g_mix = (cut * g_a + (N - cut) * g_b) / N
    assert abs(l_all - l_mix) <= TOL * abs(l_all)
    assert (g_all - g_mix).abs().max().item() <= TOL * g_all.abs().max().item()


def test_tc_loss_only_launch_matches_the_training_launch():
    """A launch without a reverse sweep (torch.no_grad) over several tiles per CTA returns the loss of the training launch
    bit for bit: the forward-only path of the two-slot W plan has to bring W_1 back for every tile by itself."""
    import pde_b200 as pb
    torch.manual_seed(3)
    for d, fn in ((3, pb.poisson.pinn_residual_loss), (4, pb.poisson.drm_energy_loss)):   # five jet channels each
        m = pb.poisson.SolutionNet(d, 64, 5, "FBC").cuda()
        N = 64 * 148 * 3 + 29
        X = torch.rand(N, d, device="cuda") * 2
        f = pb.poisson.rhs_f_for_u_sin(X, 2.0, [1] * d)
        m.zero_grad()
        l_train = fn(m, X, f, 2.0)
        l_train.backward()
        assert pb.ops.last_kernel_path() == "tcgen05"
        with torch.no_grad():
            l_eval = fn(m, X, f, 2.0)
        assert pb.ops.last_kernel_path() == "tcgen05"
        assert float(l_eval) == float(l_train.detach())


def test_path_selection():
    """Default routing: tensor-core kernel for the shapes it covers above 4096 points, generic kernel
    otherwise (fp64, wide nets); pde_set_kernel_path overrides."""
    import pde_b200 as pb
    with pb.ops.kernel_path("auto"):
        _path_selection(pb)


def _path_selection(pb):
    X = torch.rand(8192, 3, device="cuda") * 2
    f = torch.ones(8192, 1, device="cuda")
    m = pb.poisson.SolutionNet(3, 64, 5, "FBC").cuda()
    pb.poisson.pinn_residual_loss(m, X, f, 2.0)
    assert pb.ops.last_kernel_path() == "tcgen05"
    pb.poisson.pinn_residual_loss(m, X[:1000], f[:1000], 2.0)
    assert pb.ops.last_kernel_path() == "simt_fma"
    import copy
    pb.poisson.pinn_residual_loss(copy.deepcopy(m).double(), X.double(), f.double(), 2.0)
    assert pb.ops.last_kernel_path() == "simt_fma"
    m5 = pb.poisson.SolutionNet(5, 64, 5, "FBC").cuda()
    X5 = torch.rand(8192, 5, device="cuda") * 2
    pb.poisson.pinn_residual_loss(m5, X5, f, 2.0)          # 1 + 5 + 1 = 7 channels: two dimension-split tensor-core passes
    assert pb.ops.last_kernel_path() == "tcgen05"
    pb.poisson.drm_energy_loss(m5, X5, f, 2.0)             # 1 + 5 = 6 channels: tensor cores
    assert pb.ops.last_kernel_path() == "tcgen05"
    from pde_b200 import _lib as L
    from pde_b200.ops import EnvelopeSpec, ProgramSpec, residual_means
    residual_means(m, X, ProgramSpec(L.PROG_RAYLEIGH, 0.5), EnvelopeSpec(L.ENV_POLY, 0.0, 2.0), beta=f)
    assert pb.ops.last_kernel_path() == "tcgen05"          # Rayleigh quotients too (tools/rayleigh_check.py)
    m128 = pb.poisson.SolutionNet(3, 128, 5, "FBC").cuda()
    pb.poisson.pinn_residual_loss(m128, X, f, 2.0)
    assert pb.ops.last_kernel_path() == "simt_fma"
    # network jets of order <= 1 (what the WAN losses evaluate for both networks) ride the tensor-core kernel too
    J = pb.mlp_jets(m, X, 1)
    assert pb.ops.last_kernel_path() == "tcgen05"
    J.sum().backward()
    assert pb.ops.last_kernel_path() == "tcgen05"
    pb.mlp_jets(m, X, 2)
    assert pb.ops.last_kernel_path() == "simt_fma"         # Hessian-diagonal channels: generic kernel
    v = pb.poisson.CriticNet(3, 64, 3).cuda()
    Xr = X.clone().requires_grad_(True)
    lu, lv, _, _ = pb.poisson.wan_losses(m, v, Xr, f, 2.0)
    assert pb.ops.last_kernel_path() == "tcgen05"
    lv.backward()
    assert pb.ops.last_kernel_path() == "tcgen05"


@pytest.mark.parametrize("d,w,depth,order,act,N", [(3, 64, 5, 1, "sin", 777), (2, 50, 4, 1, "tanh", 4099), (1, 20, 3, 1, "tanh", 100),
                                                 (5, 24, 3, 1, "sin", 64), (2, 64, 4, 0, "sin", 1000), (4, 9, 3, 1, "sin", 333),
                                                 (4, 64, 5, 1, "sin", 64 * 148 * 2 + 100)])   # five channels, forward-only pass over >1 tile per CTA
def test_tc_jets_forward_backward_vs_numpy_oracle(d, w, depth, order, act, N):
    """pde_jets_forward / pde_jets_backward on the tensor-core kernel (orders 0, 1): jets and the reverse sweep with
    arbitrary cotangents vs oracle/jets_numpy (the same check test_gpu_poisson runs on the generic kernel)."""
    import pde_b200 as pb
    from oracle import jets_numpy as O
    rng = np.random.default_rng(50 + d)
    Ws, bs = _rand_net(rng, d, w, depth)
    X = rng.uniform(0, 2, (N, d)).astype(np.float32).astype(np.float64)
    Jbar = rng.normal(size=(N, 1 + order * d)).astype(np.float32).astype(np.float64)
    A = O.SIN if act == "sin" else O.TANH
    J, cache = O.mlp_jets_forward(Ws, bs, X, A, order)
    gWs, gbs = O.mlp_jets_backward(Ws, bs, cache, Jbar)
    net, lin = _seq(Ws, bs, act)
    Jg = pb.mlp_jets(net, torch.tensor(X, dtype=torch.float32, device="cuda"), order)
    assert pb.ops.last_kernel_path() == "tcgen05"
    assert np.max(np.abs(Jg.detach().double().cpu().numpy() - J)) <= 5 * TOL * max(1.0, np.abs(J).max())
    Jg.backward(torch.tensor(Jbar, dtype=torch.float32, device="cuda"))
    assert pb.ops.last_kernel_path() == "tcgen05"
    assert_grads_close(_grads(lin), (gWs, gbs), TOL, f"jets d{d} w{w} {act} order{order}")


def test_tc_poisson_wan_vs_reference_golden_and_large_batch():
    """wan_losses with both networks' jets and reverse sweeps on the tensor-core kernel: (i) the reference golden
    (d = 2, forced onto the tensor-core path), (ii) 2^16 points vs the float64 generic kernels."""
    import pde_b200 as pb
    from conftest import assert_grads_golden, assert_loss_close
    g = load_golden("poisson_wan_d2_w16")
    dtype = torch.float32

    def build(prefix, cls, *args):
        Ws, bs = net_from(g, prefix)
        m = cls(Ws[0].shape[1], Ws[0].shape[0], len(Ws), *args).double()
        lin = [x for x in m.net if isinstance(x, torch.nn.Linear)]
        with torch.no_grad():
            for l, W, b in zip(lin, Ws, bs):
                l.weight.copy_(torch.tensor(W)); l.bias.copy_(torch.tensor(b))
        return m.to("cuda", dtype), lin
    um, ul = build("u_", pb.poisson.SolutionNet, "FBC")
    vm, vl = build("v_", pb.poisson.CriticNet)
    X = torch.tensor(g["X"], dtype=dtype, device="cuda", requires_grad=True)
    f = torch.tensor(g["f"], dtype=dtype, device="cuda")
    lu, lv, weak, pn = pb.poisson.wan_losses(um, vm, X, f, float(g["L"]), v_reg_weight=float(g["v_reg_weight"]))
    assert pb.ops.last_kernel_path() == "tcgen05"
    for got, key in ((lu, "loss_u"), (lv, "loss_v"), (weak, "weak"), (pn, "phi_norm")):
        assert_loss_close(got, g, key, dtype, "tc wan")
    lu.backward(retain_graph=True)
    assert_grads_golden(_grads(ul), g, "lu_u_", dtype)
    assert_grads_golden(_grads(vl), g, "lu_v_", dtype)
    um.zero_grad(); vm.zero_grad()
    lv.backward()
    assert_grads_golden(_grads(ul), g, "lv_u_", dtype)
    assert_grads_golden(_grads(vl), g, "lv_v_", dtype)

    # (ii) n_interior = 2^16, 3-D, default-width networks: float32 tensor-core path vs float64 generic path
    import copy
    torch.manual_seed(4)
    u32 = pb.poisson.SolutionNet(3, 64, 5, "FBC").cuda(); v32 = pb.poisson.CriticNet(3, 64, 3).cuda()
    u64, v64 = copy.deepcopy(u32).double(), copy.deepcopy(v32).double()
    N = 1 << 16
    X = (torch.rand(N, 3, device="cuda") * 1.9 + 0.05).requires_grad_(True)
    f = pb.poisson.rhs_f_for_u_sin(X.detach(), 2.0, [1, 1, 1])
    res = {}
    for tag, (um, vm, Xc, fc, path) in {"tc": (u32, v32, X, f, "tc"),
                                        "ref": (u64, v64, X.detach().double().requires_grad_(True), f.double(), "simt")}.items():
        with pb.ops.kernel_path(path):
            lu, lv, weak, pn = pb.poisson.wan_losses(um, vm, Xc, fc, 2.0, v_reg_weight=0.5)
            lu.backward(retain_graph=True)
            gu = [p.grad.double().clone() for p in um.parameters()]
            um.zero_grad(); vm.zero_grad()
            lv.backward()
            gv = [p.grad.double().clone() for p in vm.parameters()]
        res[tag] = (lu.item(), lv.item(), gu, gv)
    assert abs(res["tc"][0] - res["ref"][0]) <= TOL * abs(res["ref"][0])
    assert abs(res["tc"][1] - res["ref"][1]) <= TOL * abs(res["ref"][1])
    for k in (2, 3):
        scale = max(float(t.abs().max()) for t in res["ref"][k])
        for a, b in zip(res["tc"][k], res["ref"][k]):
            assert float((a - b).abs().max()) <= TOL * scale


def test_tc_adjoint_scale_follows_the_residual():
    """The fp16 operand split of the reverse sweep works on adjoints scaled by a power of two that is re-derived per
    tile: a batch whose residual grows by 1e7 along a CTA's tile range (source term O(1) on the first tiles, O(1e7)
    later, and the other way round) must give the same loss and gradient as the float64 generic kernel."""
    import pde_b200 as pb
    from pde_b200 import _lib as L
    from pde_b200.ops import EnvelopeSpec, ProgramSpec, residual_means
    rng = np.random.default_rng(9)
    Ws, bs = _rand_net(rng, 3, 64, 5)
    n32, lin32 = _seq(Ws, bs, "sin")
    n64, lin64 = _seq(Ws, bs, "sin", torch.float64)
    N = 148 * 64 * 6                      # six tiles per CTA
    X = torch.rand(N, 3, device="cuda") * 1.9 + 0.05
    espec = EnvelopeSpec(L.ENV_POLY, 0.0, 2.0)
    tile = torch.arange(N, device="cuda") // 64
    for ramp in (1, -1):
        pos = (tile % 6).float() if ramp > 0 else (5 - tile % 6).float()
        f = torch.randn(N, device="cuda") * torch.pow(10.0, pos * 1.4)     # 1 ... 1e7 within every CTA's range
        for net in (n32, n64):
            net.zero_grad()
        l32 = residual_means(n32, X, ProgramSpec(L.PROG_PINN, -1.0), espec, f=f)[0]
        assert pb.ops.last_kernel_path() == "tcgen05"
        l32.backward()
        with pb.ops.kernel_path("simt"):
            l64 = residual_means(n64, X.double(), ProgramSpec(L.PROG_PINN, -1.0), espec, f=f.double())[0]
            l64.backward()
        assert math.isfinite(l32.item())
        assert abs(l32.item() - l64.item()) <= TOL * abs(l64.item())
        assert_grads_close(_grads(lin32), _grads(lin64), TOL, f"ramp {ramp}")


@pytest.mark.parametrize("act,bc,w,depth,N", [("sin", "FBC", 64, 5, 4099), ("sin", "RB", 40, 4, 777), ("tanh", "FBC", 64, 3, 1500)])
def test_tc_pinn_5d_split_vs_numpy_oracle(act, bc, w, depth, N):
    """5-D PINN (7 jet channels) on the tensor-core kernel: forward passes over directions 0..2 and 3..4, the pointwise
    residual kernel, two reverse passes whose gradients add up — loss and gradient vs oracle/jets_numpy (float64)."""
    import pde_b200 as pb
    from pde_b200 import _lib as L
    from pde_b200.ops import EnvelopeSpec, ProgramSpec, residual_means
    from oracle import jets_numpy as O
    rng = np.random.default_rng(77)
    Ws, bs = _rand_net(rng, 5, w, depth)
    X = rng.uniform(0.05, 1.95, (N, 5)).astype(np.float32).astype(np.float64)
    f = rng.normal(size=(N, 1)).astype(np.float32).astype(np.float64)
    A = O.SIN if act == "sin" else O.TANH
    want, gWs, gbs = O.poisson_pinn_loss(Ws, bs, X, f, 2.0, bc, act=A)
    net, lin = _seq(Ws, bs, act)
    env = EnvelopeSpec(L.ENV_POLY, 0.0, 2.0) if bc == "FBC" else pb.ops.NO_ENVELOPE
    loss = residual_means(net, torch.tensor(X, dtype=torch.float32, device="cuda"), ProgramSpec(L.PROG_PINN, -1.0), env,
                          f=torch.tensor(f, dtype=torch.float32, device="cuda"))[0]
    assert pb.ops.last_kernel_path() == "tcgen05"
    loss.backward()
    assert abs(loss.item() - want) <= TOL * abs(want), (loss.item(), want)
    assert_grads_close(_grads(lin), (gWs, gbs), TOL, f"5-D pinn {act} {bc}")
    with torch.no_grad():       # value only: no reverse passes
        l0 = residual_means(net, torch.tensor(X, dtype=torch.float32, device="cuda"), ProgramSpec(L.PROG_PINN, -1.0), env,
                            f=torch.tensor(f, dtype=torch.float32, device="cuda"))[0]
    assert abs(l0.item() - want) <= TOL * abs(want)


def test_tc_pinn_5d_split_large_batch_against_fp64_generic_kernel():
    """2^18 points, default-width network: the split tensor-core path vs the float64 generic kernel."""
    import pde_b200 as pb
    from pde_b200 import _lib as L
    from pde_b200.ops import EnvelopeSpec, ProgramSpec, residual_means
    rng = np.random.default_rng(5)
    Ws, bs = _rand_net(rng, 5, 64, 5)
    n32, lin32 = _seq(Ws, bs, "sin")
    n64, lin64 = _seq(Ws, bs, "sin", torch.float64)
    N = 1 << 18
    X32 = torch.rand(N, 5, device="cuda") * 1.9 + 0.05
    f32 = torch.randn(N, device="cuda")
    espec = EnvelopeSpec(L.ENV_POLY, 0.0, 2.0)
    l32 = residual_means(n32, X32, ProgramSpec(L.PROG_PINN, -1.0), espec, f=f32)[0]
    assert pb.ops.last_kernel_path() == "tcgen05"
    l32.backward()
    with pb.ops.kernel_path("simt"):
        l64 = residual_means(n64, X32.double(), ProgramSpec(L.PROG_PINN, -1.0), espec, f=f32.double())[0]
        l64.backward()
    assert abs(l32.item() - l64.item()) <= TOL * abs(l64.item())
    assert_grads_close(_grads(lin32), _grads(lin64), TOL, "5-D, 2^18 points")
