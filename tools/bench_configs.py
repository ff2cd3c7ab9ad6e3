"""Points/s of the other BASELINE.json configurations (device-resident inputs, CUDA events, median of 10)."""
import json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pde_b200 as pb
from pde_b200.schrodinger import qho_2d as Q, ipw_1d_wan as W


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def step(loss_fn, params):
    def f():
        for p in params:
            p.grad = None
        loss_fn().backward()
    return f


out = []
torch.manual_seed(0)
for name, d, N, method, bc in (("config1 Poisson 1-D PINN FBC N=20000", 1, 20000, "pinn", "FBC"),
                               ("config2 Poisson 3-D PINN FBC N=2^22", 3, 1 << 22, "pinn", "FBC"),
                               ("config3 Poisson 5-D Deep Ritz RB N=2^20", 5, 1 << 20, "drm", "RB")):
    m = pb.poisson.SolutionNet(d, 64, 5, bc).cuda()
    X = torch.rand(N, d, device="cuda") * 2
    f = pb.poisson.rhs_f_for_u_sin(X, 2.0, [1] * d)
    fn = pb.poisson.pinn_residual_loss if method == "pinn" else pb.poisson.drm_energy_loss
    ms = timed(step(lambda: fn(m, X, f, 2.0), list(m.parameters())))
    out.append({"config": name, "ms_per_step": ms, "points_per_s": N / ms * 1e3, "kernel_path": pb.ops.last_kernel_path()})
g = torch.linspace(-6, 6, 200, device="cuda")
xg, yg = torch.meshgrid(g, g, indexing="ij")
for tech in ("FBC", "FN"):
    m = Q.FCN([2, 50, 50, 50, 50, 1], 2, 1, tech).cuda()
    ms = timed(step(lambda: Q.PINN_loss(m, xg, yg, Q.Exact_energy(2, 1, 6.0), 6.0), list(m.parameters())))
    out.append({"config": f"config4 QHO 2-D PINN {tech} 200x200 grid", "ms_per_step": ms, "points_per_s": 40000 / ms * 1e3,
                "kernel_path": pb.ops.last_kernel_path()})
um = W.FCN([1, 50, 50, 50, 1], L=2.0, enforce_bc=True).cuda(); vm = W.FCN([1, 20, 20, 20, 1], L=2.0).cuda()
x = torch.linspace(0, 2, 1000, device="cuda").view(-1, 1)
ms = timed(step(lambda: W.WAN_loss(um, vm, x, 2, 2.0)[0], list(um.parameters()) + list(vm.parameters())))
out.append({"config": "config5 IPW 1-D WAN u[1,50,50,50,1] v[1,20,20,20,1] N=1000", "ms_per_step": ms, "points_per_s": 1000 / ms * 1e3,
            "kernel_path": "simt_fma (jets) + wan_kernel"})
for o in out:
    print(json.dumps(o))

# ---- whole epochs (loss step + Adam) as CUDA graphs: what the latency-bound configurations gain
def timed_epochs(run, reps=20, inner=10):
    run(); run()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(inner):
            run()
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / inner)
    return statistics.median(ts)


out2 = []
for name, d, N, method, bc in (("config1 Poisson 1-D PINN FBC N=20000", 1, 20000, "PINN", "FBC"),
                               ("config2 Poisson 3-D PINN FBC N=2^22", 3, 1 << 22, "PINN", "FBC"),
                               ("config3 Poisson 5-D Deep Ritz RB N=2^20", 5, 1 << 20, "DRM", "RB")):
    torch.manual_seed(0)
    m = pb.poisson.SolutionNet(d, 64, 5, bc).cuda()
    X = torch.rand(N, d, device="cuda") * 2
    f = pb.poisson.rhs_f_for_u_sin(X, 2.0, [1] * d)
    fn = pb.poisson.pinn_residual_loss if method == "PINN" else pb.poisson.drm_energy_loss
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    def eager():
        opt.zero_grad(); fn(m, X, f, 2.0).backward(); opt.step()
    ms_e = timed_epochs(eager, inner=3 if N > 1 << 20 else 10)
    tr = pb.train.FusedTrainer(m, 2.0, [1] * d, method=method, X=X, f=f, weights={"bc": 0.0})
    ms_g = timed_epochs(lambda: tr.step(), inner=3 if N > 1 << 20 else 10)
    out2.append({"config": name + " — epoch = loss step + Adam", "eager_ms": ms_e, "fused_graph_ms": ms_g,
                 "points_per_s_fused": N / ms_g * 1e3})
for tech in ("FBC",):
    m = Q.FCN([2, 50, 50, 50, 50, 1], 2, 1, tech).cuda()
    E = Q.Exact_energy(2, 1, 6.0)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    def eager():
        opt.zero_grad(); Q.PINN_loss(m, xg, yg, E, 6.0).backward(); opt.step()
    ms_e = timed_epochs(eager)
    optc = torch.optim.Adam(m.parameters(), lr=1e-3, capturable=True)
    def ep():
        optc.zero_grad(set_to_none=False); l = Q.PINN_loss(m, xg, yg, E, 6.0); l.backward(); optc.step(); return l.detach()
    g = pb.train.GraphedEpoch(ep); g()
    ms_g = timed_epochs(lambda: g())
    out2.append({"config": f"config4 QHO 2-D PINN {tech} 200x200 — epoch = loss step + Adam", "eager_ms": ms_e,
                 "graphed_ms": ms_g, "points_per_s_graphed": 40000 / ms_g * 1e3})
ou = torch.optim.Adam(um.parameters(), lr=1e-3); ov = torch.optim.Adam(vm.parameters(), lr=1e-3)
def wan_eager():
    for _ in range(5):
        ov.zero_grad(); W.WAN_loss(um, vm, x, 2, 2.0)[1].backward(inputs=list(vm.parameters())); ov.step()
    ou.zero_grad(); W.WAN_loss(um, vm, x, 2, 2.0)[0].backward(inputs=list(um.parameters())); ou.step()
ms_e = timed_epochs(wan_eager)
ouc = torch.optim.Adam(um.parameters(), lr=1e-3, capturable=True); ovc = torch.optim.Adam(vm.parameters(), lr=1e-3, capturable=True)
def wan_ep():
    Ju = pb.frozen_jets(um, x)
    for _ in range(5):
        ovc.zero_grad(set_to_none=False); W.WAN_loss(um, vm, x, 2, 2.0, u_jets=Ju)[1].backward(inputs=list(vm.parameters())); ovc.step()
    ouc.zero_grad(set_to_none=False); t = W.WAN_loss(um, vm, x, 2, 2.0)[0]; t.backward(inputs=list(um.parameters())); ouc.step()
    return t.detach()
g = pb.train.GraphedEpoch(wan_ep); g()
ms_g = timed_epochs(lambda: g())
out2.append({"config": "config5 IPW 1-D WAN epoch = 5 critic steps + 1 solution step + 6 Adam steps, N=1000", "eager_ms": ms_e,
             "graphed_frozen_jets_ms": ms_g})
for o in out2:
    print(json.dumps(o))
