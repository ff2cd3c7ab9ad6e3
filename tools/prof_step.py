"""One config-2 loss step (3-D PINN, SolutionNet(3,64,5), fp32) repeated a few times: the short command the
ncu captures wrap (`-k regex:tc_kernel -s 2 -c 1`).  Usage: python tools/prof_step.py [log2_points] [repeats]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pde_b200 as pb

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 22
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.manual_seed(0)
L = 2.0
m = pb.poisson.SolutionNet(3, 64, 5, "FBC").cuda()
X = torch.rand(1 << lg, 3, device="cuda") * L
f = pb.poisson.rhs_f_for_u_sin(X, L, [1, 1, 1])
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
ev[0].record()
for i in range(reps):
    m.zero_grad()
    pb.poisson.pinn_residual_loss(m, X, f, L).backward()
    ev[i + 1].record()
torch.cuda.synchronize()
print("path", pb.ops.last_kernel_path(), "ms per step", [round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(reps)])
