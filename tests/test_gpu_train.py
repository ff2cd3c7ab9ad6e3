"""GPU tests of the epoch pieces around the loss step (pde_b200.train):
point sampling + manufactured right-hand side, fused Adam, the CUDA-graph epoch against the
reference's loop (its nested-autograd loss + torch.optim.Adam, run on the CPU in float64 through
oracle/autograd_ref.py), and device-side best-model tracking."""
import math

import numpy as np
import pytest
import torch

import pde_b200 as pb
from oracle import autograd_ref as AR

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("dim", [1, 3, 5])
def test_sampling_and_rhs(dim, dtype):
    L, ks, n = 2.0, [1 + (i % 3) for i in range(dim)], 1 << 18
    X, u, f = pb.train.sample_points_rhs(n, dim, L, ks, dtype=dtype, seed=7, offset=3, want_u=True)
    assert X.shape == (n, dim) and X.dtype == dtype
    assert float(X.min()) >= 0.0 and float(X.max()) < L          # half-open like torch.rand * L
    Xd = X.double()
    assert abs(float(Xd.mean()) - L / 2) < 4 * (L / math.sqrt(12)) / math.sqrt(n * dim) * 1.5
    assert abs(float(Xd.var()) - L * L / 12) < 0.01
    # coordinates are independent of each other and of the neighbouring point
    if dim > 1:
        c = torch.corrcoef(Xd.T)
        assert float((c - torch.eye(dim, device=c.device, dtype=c.dtype)).abs().max()) < 0.01
    assert abs(float(torch.corrcoef(torch.stack([Xd[:-1, 0], Xd[1:, 0]]))[0, 1])) < 0.01
    # manufactured solution / rhs against the reference formulas (Poisson_ND.py:49-58)
    tol = 1e-12 if dtype == torch.float64 else 2e-6
    uw, fw = pb.poisson.exact_u_prod_sin(X, L, ks), pb.poisson.rhs_f_for_u_sin(X, L, ks)
    assert float((u - uw).abs().max()) <= tol
    assert float((f - fw).abs().max()) <= tol * float(fw.abs().max())
    # counter-based: same (seed, offset) -> same points; another offset or seed -> other points
    X2, _, _ = pb.train.sample_points_rhs(n, dim, L, ks, dtype=dtype, seed=7, offset=3)
    assert torch.equal(X, X2)
    X3, _, _ = pb.train.sample_points_rhs(n, dim, L, ks, dtype=dtype, seed=7, offset=4)
    X4, _, _ = pb.train.sample_points_rhs(n, dim, L, ks, dtype=dtype, seed=8, offset=3)
    assert not torch.equal(X, X3) and not torch.equal(X, X4)
    # evaluation at given points
    _, u5, f5 = pb.train.sample_points_rhs(0, 0, L, ks, X=X, want_u=True)
    assert torch.equal(u5, u) and torch.equal(f5, f)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_fused_adam_matches_torch_adam(dtype):
    torch.manual_seed(0)
    shapes = [(16, 3), (16,), (16, 16), (16,), (1, 16), (1,), ()]     # last: a trainable energy scalar
    ours = [torch.randn(s, dtype=dtype, device="cuda") for s in shapes]
    theirs = [p.clone().requires_grad_(True) for p in ours]
    opt_t = torch.optim.Adam(theirs, lr=1e-3)
    opt = pb.train.FusedAdam(ours, lr=1e-3)
    n = sum(p.numel() for p in ours)
    for it in range(12):
        g = torch.randn(n + 3, dtype=dtype, device="cuda") * (10.0 ** (it % 3 - 1))
        o = 0
        for p in theirs:
            p.grad = (0.5 * g[o:o + p.numel()]).view_as(p).clone(); o += p.numel()
        opt_t.step()
        opt.step(g, grad_scale=0.5)
    assert int(opt.step_count) == 12
    tol = 1e-12 if dtype == torch.float64 else 2e-6
    for a, b in zip(ours, theirs):
        assert float((a - b.detach()).abs().max()) <= tol * max(1.0, float(b.detach().abs().max()))


def _reference_loop(model64_cpu, X, f, L, bc, method, lr, epochs):
    """train_poisson_nd's PINN / DRM branch (Poisson_ND.py:215-241) restated with the oracle's
    nested-autograd loss: loss -> backward -> torch.optim.Adam.step(), float64 on the CPU."""
    net = model64_cpu.net
    opt = torch.optim.Adam(net.parameters(), lr=lr)
    losses = []
    for _ in range(epochs):
        loss, gflat = AR.loss_and_grads(method, net, X, f, L, bc)
        o = 0
        for p in net.parameters():
            p.grad = gflat[o:o + p.numel()].view_as(p).clone(); o += p.numel()
        opt.step()
        losses.append(loss)
    return losses


@pytest.mark.parametrize("method,bc,dim", [("PINN", "FBC", 2), ("DRM", "RB", 3), ("PINN", "FBC", 3)])
@pytest.mark.parametrize("graph", [True, False])
def test_trainer_matches_reference_loop(method, bc, dim, graph):
    torch.manual_seed(11)
    L, ks, N, lr, epochs = 2.0, [1] * dim, 4096 + 37, 1e-3, 6
    m_cpu = pb.poisson.SolutionNet(dim, 32, 4, bc).double()
    m_gpu = pb.poisson.SolutionNet(dim, 32, 4, bc).double()
    m_gpu.load_state_dict(m_cpu.state_dict())
    m_gpu = m_gpu.cuda()
    X = torch.rand(N, dim, dtype=torch.float64) * L
    f = pb.poisson.rhs_f_for_u_sin(X, L, ks)
    want = _reference_loop(m_cpu, X, f, L, bc, method.lower(), lr, epochs)
    tr = pb.train.FusedTrainer(m_gpu, L, ks, method=method, lr=lr, X=X.cuda(), f=f.cuda(), history=epochs, graph=graph)
    tr.step(epochs)
    got = tr.hist_loss.cpu().numpy()
    np.testing.assert_allclose(got, np.array(want), rtol=1e-9, atol=1e-12)
    for a, b in zip(m_gpu.parameters(), m_cpu.parameters()):
        assert float((a.detach().cpu() - b.detach()).abs().max()) <= 1e-9 * max(1.0, float(b.detach().abs().max()))
    assert int(tr.opt.step_count) == epochs and tr.epochs_done == epochs


def test_trainer_fp32_tensor_core_path_and_best_tracking():
    """Config-2 shaped network on the tcgen05 path: graph replays are deterministic, the L2 history's minimum
    is what keep_best kept, and load_best restores those parameters."""
    torch.manual_seed(5)
    L, ks = 2.0, [1, 1, 1]
    def run():
        torch.manual_seed(5)
        m = pb.poisson.SolutionNet(3, 64, 5, "FBC").cuda()
        tr = pb.train.FusedTrainer(m, L, ks, method="PINN", n_interior=1 << 14, lr=2e-3, n_test=4096, history=20, seed=3)
        tr.step(20)
        return m, tr
    m1, t1 = run()
    m2, t2 = run()
    for a, b in zip(m1.parameters(), m2.parameters()):
        assert torch.equal(a, b)
    h = t1.hist_l2sq.double().cpu().numpy()
    assert h.min() > 0 and abs(float(t1.best_metric) / 4096 - h.min()) <= 1e-6 * h.min()
    assert int(t1.best_step) == int(np.argmin(h)) + 1
    assert abs(float(t1.l2) - math.sqrt(h[-1])) <= 1e-6
    losses = t1.hist_loss.cpu().numpy()
    assert losses[-1] < losses[0]                      # Adam makes progress on the PINN residual
    # the kept parameters reproduce the best metric on the same test draw
    t1.load_best()
    Xt, ut, _ = pb.train.sample_points_rhs(4096, 3, L, ks, seed=3 ^ 0x9E3779B97F4A7C15, offset=int(t1.best_step), want_u=True)
    with torch.no_grad():
        l2sq = float(((m1(Xt, L) - ut) ** 2).mean())
    assert abs(l2sq - h.min()) <= 2e-5 * h.min()


def test_trainer_resamples_every_epoch():
    torch.manual_seed(1)
    m = pb.poisson.SolutionNet(2, 32, 4, "FBC").cuda()
    tr = pb.train.FusedTrainer(m, 2.0, [1, 2], method="DRM", n_interior=8192, resample=True, seed=9, history=4)
    seen = []
    for _ in range(4):
        tr.step()
        seen.append(tr.X.clone())
    assert not torch.equal(seen[0], seen[1]) and not torch.equal(seen[2], seen[3])
    # epoch e draws the points of Philox offset e (the step counter before the update)
    for e in (0, 1, 3):
        Xe, _, fe = pb.train.sample_points_rhs(8192, 2, 2.0, [1, 2], seed=9, offset=e)
        assert torch.equal(seen[e], Xe)
    assert torch.isfinite(tr.hist_loss).all()


def test_wan_frozen_jets_give_identical_losses_and_grads():
    """§8f-2: the critic steps reuse the frozen u-network's jets on fixed points (IPW_1D_WAN.py:186-194)."""
    from pde_b200.schrodinger import ipw_1d_wan as W
    torch.manual_seed(2)
    L, n = 2.0, 2
    um = W.FCN([1, 50, 50, 50, 1], L=L, enforce_bc=True).cuda()
    vm = W.FCN([1, 20, 20, 20, 1], L=L).cuda()
    x = torch.linspace(0, L, 1000, device="cuda").view(-1, 1).requires_grad_(True)
    for p in um.parameters():
        p.requires_grad_(False)
    _, lv_a, lp_a, _ = W.WAN_loss(um, vm, x, n, L)
    lv_a.backward()
    ga = [p.grad.clone() for p in vm.parameters()]
    for p in vm.parameters():
        p.grad = None
    Ju = pb.frozen_jets(um, x)
    _, lv_b, lp_b, _ = W.WAN_loss(um, vm, x, n, L, u_jets=Ju)
    lv_b.backward()
    assert torch.equal(lv_a, lv_b) and torch.equal(lp_a, lp_b)
    for a, p in zip(ga, vm.parameters()):
        assert torch.equal(a, p.grad)
    assert all(p.grad is None for p in um.parameters())
    with pytest.raises(ValueError):
        W.WAN_loss(um, vm, x, n, L, u_jets=Ju[:10])


def test_graphed_wan_epoch_matches_eager():
    """One epoch of the IPW 1-D WAN loop (5 critic steps + 1 solution step, IPW_1D_WAN.py:186-208) captured as a
    CUDA graph equals the same epoch run eagerly."""
    from pde_b200.schrodinger import ipw_1d_wan as W
    L, n = 2.0, 2
    x = torch.linspace(0, L, 1000, device="cuda").view(-1, 1)

    def build():
        torch.manual_seed(4)
        um = W.FCN([1, 50, 50, 50, 1], L=L, enforce_bc=True).cuda()
        vm = W.FCN([1, 20, 20, 20, 1], L=L).cuda()
        ou = torch.optim.Adam(um.parameters(), lr=1e-3, capturable=True)
        ov = torch.optim.Adam(vm.parameters(), lr=1e-3, capturable=True)

        def epoch():
            Ju = pb.frozen_jets(um, x)
            for _ in range(5):
                ov.zero_grad(set_to_none=False)
                _, lv, _, _ = W.WAN_loss(um, vm, x, n, L, u_jets=Ju)
                gv = torch.autograd.grad(lv, list(vm.parameters()))
                for p, g in zip(vm.parameters(), gv):
                    p.grad = g if p.grad is None else p.grad.copy_(g)
                ov.step()
            ou.zero_grad(set_to_none=False)
            total, _, lp, ln = W.WAN_loss(um, vm, x, n, L)
            gu = torch.autograd.grad(total, list(um.parameters()))
            for p, g in zip(um.parameters(), gu):
                p.grad = g if p.grad is None else p.grad.copy_(g)
            ou.step()
            return total.detach()
        return um, vm, epoch

    um1, vm1, ep1 = build()
    for _ in range(6):
        t1 = ep1()
    um2, vm2, ep2 = build()
    g = pb.train.GraphedEpoch(ep2, warmup=3)
    g()                      # 3 eager epochs + capture
    for _ in range(3):
        t2 = g()             # 3 replays
    torch.cuda.synchronize()
    assert g.replays == 3
    for a, b in zip(list(um1.parameters()) + list(vm1.parameters()), list(um2.parameters()) + list(vm2.parameters())):
        assert float((a - b).abs().max()) <= 1e-6 * max(1.0, float(a.abs().max()))
    assert abs(float(t1) - float(t2)) <= 1e-5 * max(1e-3, abs(float(t1)))
