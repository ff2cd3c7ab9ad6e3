"""One config-3 loss step (5-D Deep Ritz, SolutionNet(5,64,5,'RB'), 2^20 points, fp32) repeated: the command the ncu
capture / timeline of the 6-channel (W-streamed) variant wraps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pde_b200 as pb
torch.manual_seed(0)
m = pb.poisson.SolutionNet(5, 64, 5, "RB").cuda()
X = torch.rand(1 << 20, 5, device="cuda") * 2
f = pb.poisson.rhs_f_for_u_sin(X, 2.0, [1] * 5)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
ev[0].record()
for i in range(3):
    m.zero_grad()
    pb.poisson.drm_energy_loss(m, X, f, 2.0).backward()
    ev[i + 1].record()
torch.cuda.synchronize()
print("path", pb.ops.last_kernel_path(), "ms per step", [round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(3)])
