"""Config 4's fused epoch (QHO_2D eigenstate PINN on the 200 x 200 grid + Adam, replayed as one CUDA graph), repeated:
the command the ncu launch list of that epoch wraps.  Usage: python tools/prof_cfg4_epoch.py [replays]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pde_b200 as pb
from pde_b200.schrodinger import qho_2d as Q

torch.manual_seed(0)
g1 = torch.linspace(-6.0, 6.0, 200)
xg, yg = torch.meshgrid(g1, g1, indexing="ij")
m = Q.FCN([2, 50, 50, 50, 50, 1], 2, 1, "FBC").cuda()
xd, yd = xg.cuda(), yg.cuda()
E = Q.Exact_energy(2, 1, 6.0)
opt = pb.train.Adam(m.parameters(), lr=1e-3)


def epoch():
    opt.zero_grad(set_to_none=False)
    l = Q.PINN_loss(m, xd, yd, E, 6.0); l.backward(); opt.step()
    return l.detach()


ge = pb.train.GraphedEpoch(epoch); ge()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    ge()
e1.record(); torch.cuda.synchronize()
print(f"config 4 fused epoch: {e0.elapsed_time(e1) / n:.4f} ms ({pb.ops.last_kernel_path()})")
