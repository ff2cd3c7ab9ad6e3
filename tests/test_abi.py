"""CPU tests of the drop-in boundary: libpde_b200.so loads, exports every symbol that
include/pde_b200.h declares, and its pure (no-GPU) queries answer; the ctypes structures match
the header's layout.  No compute call is made here."""
import ctypes as C
import os
import re

import pytest

import pde_b200 as pb
from pde_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "pde_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pde_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_are_exported():
    lib = pb.load_library()
    names = _declared_functions()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pde_b200.h but not exported"
    assert sorted(L.EXPORTS) == names, "ctypes binding and header disagree on the entry points"


def test_pure_queries():
    lib = pb.load_library()
    assert lib.pde_abi_version() == 1
    assert lib.pde_strerror(0) == b"ok"
    assert b"workspace" in lib.pde_strerror(-3)
    assert lib.pde_jet_channels(3, 2) == 7 and lib.pde_jet_channels(5, 1) == 6 and lib.pde_jet_channels(2, 0) == 1
    assert lib.pde_jet_channels(6, 2) < 0
    assert [lib.pde_program_quantities(k) for k in (1, 2, 3, 4)] == [1, 1, 2, 1]
    assert [lib.pde_program_order(k) for k in (1, 2, 3, 4)] == [2, 1, 1, 0]
    assert lib.pde_program_order(99) < 0


def test_param_count_and_validation():
    lib = pb.load_library()
    net = L.Net()
    net.dtype, net.dim, net.n_linear, net.activation = L.F32, 3, 5, L.ACT_SIN
    for i, w in enumerate([3, 64, 64, 64, 64, 1]):
        net.widths[i] = w
    n = C.c_int64(0)
    assert lib.pde_param_count(C.byref(net), C.byref(n)) == 0
    assert n.value == 12801          # SolutionNet(3, 64, 5): Poisson_ND.py:13-23
    net.widths[2] = 32               # unequal hidden widths are not a reference shape
    assert lib.pde_param_count(C.byref(net), C.byref(n)) == -2
    net.widths[2] = 64
    net.dim = 9
    assert lib.pde_param_count(C.byref(net), C.byref(n)) == -2


def test_struct_layout_matches_header():
    # sizes follow from the C declarations (int32 / double / pointer members, natural alignment)
    assert C.sizeof(L.Net) == 4 * 4 + 9 * 4 + 4 + 8 * 8 + 8 * 8        # 52 -> padded to 56, then 16 pointers
    assert C.sizeof(L.Envelope) == 4 + 5 * 4 + 2 * 8 + 5 * 8 * 8
    assert C.sizeof(L.Program) == 8 + 3 * 8 + 3 * 8
    assert L.Wan.env_u.offset == 8 + 6 * 8 + 3 * 8


def test_no_cpu_fallback():
    """Host tensors are refused loudly; nothing routes through the oracle."""
    import torch
    m = pb.poisson.SolutionNet(2, 16, 3)
    X = torch.rand(8, 2) * 2
    with pytest.raises(pb.PdeError):
        pb.poisson.pinn_residual_loss(m, X, torch.ones(8, 1), 2.0)
    src = open(os.path.join(ROOT, "neural-network-based-pde-solver_b200", "ops.py")).read()
    src += open(os.path.join(ROOT, "neural-network-based-pde-solver_b200", "poisson.py")).read()
    assert "oracle" not in src


def test_host_side_settings():
    """Process-wide settings and counters that need no device: kernel-family override, exchange timeout, launch counter."""
    lib = pb.load_library()
    old = lib.pde_kernel_path()
    try:
        for v in (-1, 0, 1):
            assert lib.pde_set_kernel_path(v) == 0 and lib.pde_kernel_path() == v
        assert lib.pde_set_kernel_path(2) == -1 and lib.pde_set_kernel_path(-3) == -1
        with pb.ops.kernel_path("simt"):
            assert lib.pde_kernel_path() == 0
            with pb.ops.kernel_path("tc"):
                assert lib.pde_kernel_path() == 1
            assert lib.pde_kernel_path() == 0
    finally:
        lib.pde_set_kernel_path(old)
    assert lib.pde_set_exchange_timeout(600.0) == 0 and lib.pde_set_exchange_timeout(0.0) == 0
    assert lib.pde_set_exchange_timeout(-1.0) == -1 and lib.pde_set_exchange_timeout(float("nan")) == -1
    lib.pde_set_exchange_timeout(600.0)
    assert lib.pde_last_kernel_path() in (-1, 0, 1)
    assert isinstance(pb.ops.launch_count(), int) and pb.ops.launch_count() >= 0
    # invalid arguments are refused before anything is enqueued
    assert lib.pde_wan_scalars(L.F32, 5, None, None, None, None, None) == -1
    assert lib.pde_residual_loss_grad_exchange(None, None, None, None, 0, None, 1.0, None, None, 0, None, 0, None, None) == -1


def test_train_adam_takes_adopted_gradients_without_copy(monkeypatch):
    """Host logic of pb.train.Adam (the launch itself needs a GPU and is covered by tests/test_gpu_train.py): after
    zero_grad() autograd adopts the views of one flat gradient buffer that the drop-in operators return, and step()
    hands that buffer to the kernel as it is; any other layout is gathered; set_to_none=False keeps the optimiser's
    own views."""
    import torch
    from pde_b200 import train as T

    seen = []

    class Recorder:           # stands in for FusedAdam (which insists on CUDA tensors)
        def __init__(self, params, **kw):
            self.n = sum(p.numel() for p in params)
            self.exp_avg = self.exp_avg_sq = self.step_count = None

        def step(self, flat, grad_scale=1.0):
            seen.append(flat)

    monkeypatch.setattr(T, "FusedAdam", Recorder)

    class FlatGrad(torch.autograd.Function):      # the shape of ops._Residual.backward / ops._Jets.backward
        @staticmethod
        def forward(ctx, *ps):
            ctx.shapes = [p.shape for p in ps]
            return sum((p * p).sum() for p in ps).detach() * 0.5

        @staticmethod
        def backward(ctx, g):
            n = sum(s.numel() for s in ctx.shapes)
            flat = torch.arange(1, n + 1, dtype=torch.float64) * g
            FlatGrad.last = flat
            out, o = [], 0
            for s in ctx.shapes:
                out.append(flat[o:o + s.numel()].view(s)); o += s.numel()
            return tuple(out)

    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Linear(3, 8), torch.nn.Tanh(), torch.nn.Linear(8, 1)).double()
    ps = list(m.parameters())
    n = sum(p.numel() for p in ps)
    want = torch.arange(1, n + 1, dtype=torch.float64)
    opt = T.Adam(ps, lr=1e-3)

    # adopted: no copy, the kernel would read the operator's own buffer
    opt.zero_grad()
    FlatGrad.apply(*ps).backward()
    opt.step()
    assert opt.adopted_steps == 1 and opt.gathered_steps == 0
    assert seen[-1].data_ptr() == FlatGrad.last.data_ptr() and torch.equal(seen[-1], want)
    # a second backward() accumulates into the adopted views in place: layout kept
    opt.zero_grad()
    FlatGrad.apply(*ps).backward()
    (2.0 * FlatGrad.apply(*ps)).backward()
    opt.step()
    assert opt.adopted_steps == 2 and opt.gathered_steps == 0 and torch.equal(seen[-1], 3.0 * want)
    # two terms in one backward(): the engine sums them per parameter before they reach p.grad -> gathered, same values
    opt.zero_grad()
    (FlatGrad.apply(*ps) + 2.0 * FlatGrad.apply(*ps)).backward()
    opt.step()
    assert opt.adopted_steps == 2 and opt.gathered_steps == 1 and torch.equal(seen[-1], 3.0 * want)
    opt.gathered_steps = 0
    # gradients of ordinary torch operators: gathered, same values as torch.cat
    opt.zero_grad()
    m(torch.ones(5, 3, dtype=torch.float64)).sum().backward()
    ref = torch.cat([p.grad.reshape(-1) for p in ps])
    opt.step()
    assert opt.gathered_steps == 1 and torch.equal(seen[-1], ref)
    # a parameter without gradient counts as zero; a subset (backward(inputs=...)) is gathered
    opt.zero_grad()
    FlatGrad.apply(*ps).backward(inputs=ps[:2])
    opt.step()
    k = ps[0].numel() + ps[1].numel()
    assert opt.gathered_steps == 2 and torch.equal(seen[-1][:k], want[:k]) and not seen[-1][k:].any()
    # set_to_none=False: the optimiser's own views, zeroed by one fill, accumulated into by autograd
    opt.zero_grad(set_to_none=False)
    assert all(p.grad.data_ptr() == opt.flat.data_ptr() + off * 8 for p, off in zip(ps, opt._off))
    FlatGrad.apply(*ps).backward()
    opt.step()
    assert seen[-1].data_ptr() == opt.flat.data_ptr() and torch.equal(seen[-1], want)
    opt.zero_grad(set_to_none=False)
    assert not opt.flat.any() and all(p.grad.data_ptr() == opt.flat.data_ptr() + off * 8 for p, off in zip(ps, opt._off))
    # a different parameter order than the operator's: not consecutive, hence gathered — never misread
    opt2 = T.Adam(ps[::-1], lr=1e-3)
    opt2.zero_grad()
    FlatGrad.apply(*ps).backward()
    opt2.step()
    assert opt2.gathered_steps == 1 and torch.equal(seen[-1], torch.cat([p.grad.reshape(-1) for p in ps[::-1]]))


def test_linear_stack_reads_every_reference_module_layout():
    """ops._linear_stack finds the Linear layers and the activation in each layout the reference's scripts use
    (Sequential under .net; UnifiedEigenModel.u_model.net; FCN_Single.net -> FCN.layers ModuleList with the
    activation as an attribute; a bare Sequential) and refuses what the kernels do not implement."""
    import torch
    import torch.nn as nn
    from pde_b200 import ops
    from pde_b200.schrodinger import ipw_1d_wan, kh_1d, qho_1d_pinn_drm, qho_2d

    def widths(lin):
        return [lin[0].in_features] + [m.out_features for m in lin]

    lin, act = ops._linear_stack(pb.poisson.SolutionNet(3, 64, 5, "FBC"))
    assert widths(lin) == [3, 64, 64, 64, 64, 1] and act == "sin"
    lin, act = ops._linear_stack(pb.poisson.CriticNet(2, 16, 3))
    assert widths(lin) == [2, 16, 16, 1] and act == "sin"
    lin, act = ops._linear_stack(ipw_1d_wan.FCN([1, 50, 50, 50, 1], L=2.0, enforce_bc=True))
    assert widths(lin) == [1, 50, 50, 50, 1] and act == "tanh"
    lin, act = ops._linear_stack(qho_2d.FCN([2, 50, 50, 50, 50, 1], 2, 1, "FN"))
    assert widths(lin) == [2, 50, 50, 50, 50, 1] and act == "sin"
    um = kh_1d.UnifiedEigenModel([1, 100, 100, 100, 1], "FBC", 60.0)
    lin, act = ops._linear_stack(um)
    assert widths(lin) == [1, 100, 100, 100, 1] and lin[0] is um.u_model.net[0]
    single = qho_1d_pinn_drm.FCN_Single([1, 20, 20, 1], num_states=1)
    lin, act = ops._linear_stack(single)
    assert widths(lin) == [1, 20, 20, 1] and act in ("tanh", "sin")
    lin, act = ops._linear_stack(nn.Sequential(nn.Linear(2, 8), nn.Tanh(), nn.Linear(8, 1)))
    assert widths(lin) == [2, 8, 1] and act == "tanh"
    # the parameter order the kernels' flat gradient uses is nn.Module.parameters() order
    m = pb.poisson.SolutionNet(2, 8, 3, "RB")
    lin, _ = ops._linear_stack(m)
    assert [id(p) for l in lin for p in (l.weight, l.bias)] == [id(p) for p in m.parameters()]
    with pytest.raises(NotImplementedError):
        ops._linear_stack(nn.Sequential(nn.Linear(2, 8), nn.ReLU(), nn.Linear(8, 1)))
    with pytest.raises(NotImplementedError):
        ops._linear_stack(nn.Sequential(nn.Linear(2, 8), nn.Tanh(), nn.Linear(8, 8), pb.poisson.Sin(), nn.Linear(8, 1)))
    with pytest.raises(ValueError):
        ops._linear_stack(nn.Sequential(nn.Tanh()))
    # points on the CPU: refused loudly, never computed some other way
    with pytest.raises(L.PdeError):
        ops._Net(m, torch.zeros(4, 2))


def test_envelope_struct_cache_follows_the_description():
    """EnvelopeSpec.to_c caches the ctypes struct per description and rebuilds it when a field changes."""
    from pde_b200.ops import EnvelopeSpec
    e = EnvelopeSpec(L.ENV_POLY, 0.0, 2.0, [[0.5, 1.5]])
    c1 = e.to_c()
    assert e.to_c() is c1 and c1.kind == L.ENV_POLY and c1.hi == 2.0 and c1.n_nodes[0] == 2 and c1.nodes[0][1] == 1.5
    e.hi = 3.0
    c2 = e.to_c()
    assert c2 is not c1 and c2.hi == 3.0 and e.to_c() is c2
    with pytest.raises(NotImplementedError):
        EnvelopeSpec(L.ENV_POLY, 0.0, 2.0, [[0.1 * k for k in range(L.MAX_NODES + 1)]]).to_c()
