"""The graphed epochs bench.py reports for the latency-bound configurations, replayed a few times each: the command
the ncu launch list of those epochs wraps (profiles/r2h_graphed_epochs_launches.csv).

  config 4   QHO_2D eigenstate PINN on the 200 x 200 grid + pde_b200.train.Adam
  config 5   IPW_1D_WAN minimax epoch (five critic updates on the frozen solution network's jets + one solution update)

Usage: python tools/prof_graphed_epochs.py [replays] [--own-views] [--tc]
       (--own-views: zero_grad(set_to_none=False); --tc: tensor-core kernels also below their 4096-point threshold)
Epoch boundaries in the launch list follow from the kernel names: one tc_kernel per config-4 epoch, six
wan_scalars_kernel per config-5 epoch (tools/epoch_launch_counts.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pde_b200 as pb
from pde_b200.schrodinger import ipw_1d_wan as W
from pde_b200.schrodinger import qho_2d as Q

args = [a for a in sys.argv[1:] if not a.startswith("--")]
n = int(args[0]) if args else 3
none = "--own-views" not in sys.argv
if "--tc" in sys.argv:
    pb.ops.kernel_path("tc").__enter__()


def timed(ge):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        ge()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


torch.manual_seed(0)
g1 = torch.linspace(-6.0, 6.0, 200)
xg, yg = torch.meshgrid(g1, g1, indexing="ij")
m = Q.FCN([2, 50, 50, 50, 50, 1], 2, 1, "FBC").cuda()
xd, yd = xg.cuda(), yg.cuda()
E = Q.Exact_energy(2, 1, 6.0)
opt = pb.train.Adam(m.parameters(), lr=1e-3)


def q_epoch():
    opt.zero_grad(set_to_none=none)
    l = Q.PINN_loss(m, xd, yd, E, 6.0); l.backward(); opt.step()
    return l.detach()


ge = pb.train.GraphedEpoch(q_epoch); ge()
print(f"config 4 graphed epoch: {timed(ge):.4f} ms ({pb.ops.last_kernel_path()}; adopted {opt.adopted_steps}, gathered {opt.gathered_steps})")

um = W.FCN([1, 50, 50, 50, 1], L=2.0, enforce_bc=True).cuda()
vm = W.FCN([1, 20, 20, 20, 1], L=2.0).cuda()
x = torch.linspace(0, 2, 1000, device="cuda").view(-1, 1)
ou, ov = pb.train.Adam(um.parameters(), lr=1e-3), pb.train.Adam(vm.parameters(), lr=1e-3)


def wan_epoch():     # IPW_1D_WAN.py:186-208
    Ju = pb.frozen_jets(um, x)
    for _ in range(5):
        ov.zero_grad(set_to_none=none)
        W.WAN_loss(um, vm, x, 2, 2.0, u_jets=Ju)[1].backward(inputs=list(vm.parameters())); ov.step()
    ou.zero_grad(set_to_none=none)
    t = W.WAN_loss(um, vm, x, 2, 2.0)[0]; t.backward(inputs=list(um.parameters())); ou.step()
    return t.detach()


gw = pb.train.GraphedEpoch(wan_epoch); gw()
print(f"config 5 graphed minimax epoch: {timed(gw):.4f} ms (adopted {ou.adopted_steps + ov.adopted_steps}, "
      f"gathered {ou.gathered_steps + ov.gathered_steps})")
