"""Development check (GPU): tcgen05 path vs the numpy oracle, per-tensor errors printed.
    the kernel family is set per case (pde_set_kernel_path); run as `python tools/tc_check.py [quick]`."""
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import pde_b200 as pb
from pde_b200 import _lib as L
from pde_b200.ops import EnvelopeSpec, ProgramSpec, residual_means
from oracle import jets_numpy as O


def make_net(d, w, depth, act, rng, dtype=torch.float32):
    Ws = [rng.uniform(-1, 1, (w, d)) / math.sqrt(d)] + [rng.uniform(-1, 1, (w, w)) / math.sqrt(w) for _ in range(depth - 2)] \
        + [rng.uniform(-1, 1, (1, w)) / math.sqrt(w)]
    bs = [rng.uniform(-0.5, 0.5, W.shape[0]) for W in Ws]
    # the kernel sees fp32 weights: round them first so the oracle uses identical values
    Ws = [W.astype(np.float32).astype(np.float64) for W in Ws]
    bs = [b.astype(np.float32).astype(np.float64) for b in bs]
    mods = []
    for i in range(depth - 1):
        mods += [torch.nn.Linear(Ws[i].shape[1], w), pb.poisson.Sin() if act == "sin" else torch.nn.Tanh()]
    mods += [torch.nn.Linear(w, 1)]
    net = torch.nn.Sequential(*mods).double()
    lin = [x for x in net if isinstance(x, torch.nn.Linear)]
    with torch.no_grad():
        for l, W, b in zip(lin, Ws, bs):
            l.weight.copy_(torch.tensor(W)); l.bias.copy_(torch.tensor(b))
    return net.to("cuda", dtype), lin, Ws, bs


def report(tag, loss, want, lin, gWs, gbs):
    worst = 0.0
    scale = max(max(np.max(np.abs(w)) for w in gWs), max(np.max(np.abs(b)) for b in gbs))
    rows = []
    for i, (l, gW, gb) in enumerate(zip(lin, gWs, gbs)):
        for nm, a, b in (("W", l.weight.grad, gW), ("b", l.bias.grad, gb)):
            a = a.double().cpu().numpy().reshape(b.shape)
            rl2 = np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)
            rmx = np.max(np.abs(a - b)) / scale
            rows.append(f"{nm}{i}:{rl2:.1e}/{rmx:.1e}")
            worst = max(worst, rmx, rl2 if np.linalg.norm(b) > 1e-3 * scale * math.sqrt(b.size) else 0.0)
    lerr = abs(loss - want) / max(abs(want), 1e-3)
    ok = lerr <= 1e-5 and worst <= 1e-5
    print(f"[{'OK ' if ok else 'BAD'}] {tag}: loss rel {lerr:.2e} ({loss:.6g} vs {want:.6g}) worst grad {worst:.2e} | " + " ".join(rows), flush=True)
    return ok


def case(d, w, depth, act, prog, env_kind, N, seed=0, path="tc"):
    L.load().pde_set_kernel_path({"simt": 0, "tc": 1}[path])
    rng = np.random.default_rng(seed)
    net, lin, Ws, bs = make_net(d, w, depth, act, rng)
    X = rng.uniform(0.05, 1.95, (N, d)).astype(np.float32).astype(np.float64)
    f = rng.normal(size=(N, 1)).astype(np.float32).astype(np.float64)
    beta = rng.uniform(0.5, 1.5, (N, 1)).astype(np.float32).astype(np.float64)
    A = O.SIN if act == "sin" else O.TANH
    env = {"kind": {"poly": O.ENV_POLY, "exp": O.ENV_EXPWIN, "none": O.ENV_NONE}[env_kind], "lo": 0.0, "hi": 2.0}
    espec = EnvelopeSpec({"poly": L.ENV_POLY, "exp": L.ENV_EXPWIN, "none": L.ENV_NONE}[env_kind], 0.0, 2.0)
    Xg = torch.tensor(X, dtype=torch.float32, device="cuda")
    fg = torch.tensor(f, dtype=torch.float32, device="cuda")
    bg = torch.tensor(beta, dtype=torch.float32, device="cuda")
    if prog == "pinn":
        want, gWs, gbs, _ = O.eigen_pinn_loss(Ws, bs, X, A, env, -1.0, beta, 0.3, f=f)
        m = residual_means(net, Xg, ProgramSpec(L.PROG_PINN, -1.0, 0.0, 0.3), espec, f=fg, beta=bg)
        loss = m[0]
    elif prog == "drm":
        def program(U):
            qq, Ubar = O.drm_poisson_program(U, d, f)
            return float(qq.mean()), Ubar / N, None
        want, gWs, gbs, _ = O._loss_and_grads(Ws, bs, X, A, 1, env, program)
        m = residual_means(net, Xg, ProgramSpec(L.PROG_DRM, 0.5), espec, f=fg)
        loss = m[0]
    elif prog == "rayleigh":
        want, gWs, gbs, _ = O.rayleigh_loss(Ws, bs, X, A, env, 0.5, beta)
        m = residual_means(net, Xg, ProgramSpec(L.PROG_RAYLEIGH, 0.5), espec, beta=bg)
        loss = m[0] / m[1]
    else:
        raise ValueError(prog)
    loss.backward()
    torch.cuda.synchronize()
    return report(f"{path} d{d} w{w} depth{depth} {act} {prog} env={env_kind} N={N}", float(loss.item()), want, lin, gWs, gbs)


def big_case(N=1 << 20, d=3, seed=3):
    """tcgen05 fp32 vs the generic kernel in fp64 on the same large batch (accumulation accuracy)."""
    rng = np.random.default_rng(seed)
    net32, lin32, Ws, bs = make_net(d, 64, 5, "sin", rng)
    net64, lin64, _, _ = make_net(d, 64, 5, "sin", np.random.default_rng(seed), dtype=torch.float64)
    X = torch.rand(N, d, device="cuda", dtype=torch.float64) * 1.9 + 0.05
    f = torch.randn(N, device="cuda", dtype=torch.float64)
    X32 = X.float(); f32 = f.float()
    X = X32.double(); f = f32.double()
    espec = EnvelopeSpec(L.ENV_POLY, 0.0, 2.0)
    L.load().pde_set_kernel_path(1)
    l32 = residual_means(net32, X32, ProgramSpec(L.PROG_PINN, -1.0), espec, f=f32)[0]
    l32.backward()
    L.load().pde_set_kernel_path(0)
    l64 = residual_means(net64, X, ProgramSpec(L.PROG_PINN, -1.0), espec, f=f)[0]
    l64.backward()
    torch.cuda.synchronize()
    gWs = [l.weight.grad.cpu().numpy() for l in lin64]; gbs = [l.bias.grad.cpu().numpy() for l in lin64]
    return report(f"tc fp32 vs generic fp64, d{d} N={N}", float(l32.item()), float(l64.item()), lin32, gWs, gbs)


def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    ok = True
    ok &= case(3, 64, 5, "sin", "pinn", "poly", 64)
    if quick:
        return 0 if ok else 1
    ok &= case(3, 64, 5, "sin", "pinn", "poly", 1000)
    ok &= case(3, 64, 5, "sin", "pinn", "poly", 64 * 148 * 2 + 17)
    ok &= case(5, 64, 5, "sin", "drm", "none", 3000)
    ok &= case(1, 64, 5, "sin", "pinn", "poly", 777)
    ok &= case(2, 50, 5, "sin", "pinn", "exp", 2049)
    ok &= case(1, 50, 4, "tanh", "pinn", "poly", 1000)
    ok &= case(2, 50, 5, "sin", "rayleigh", "poly", 1500)
    ok &= case(4, 33, 3, "tanh", "pinn", "none", 500)
    ok &= case(3, 64, 5, "sin", "pinn", "poly", 1000, path="simt")
    ok &= big_case()
    print("ALL OK" if ok else "SOME BAD")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
