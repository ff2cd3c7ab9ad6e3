"""One IPW_1D_WAN loss evaluation + backward at the reference's size (1000 points, u [1,50,50,50,1], v [1,20,20,20,1]),
repeated: the command the ncu launch list of the latency-bound WAN path wraps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pde_b200 as pb
from pde_b200.schrodinger import ipw_1d_wan as W
torch.manual_seed(0)
um = W.FCN([1, 50, 50, 50, 1], L=2.0, enforce_bc=True).cuda(); vm = W.FCN([1, 20, 20, 20, 1], L=2.0).cuda()
x = torch.linspace(0, 2, 1000, device="cuda").view(-1, 1)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    for p in list(um.parameters()) + list(vm.parameters()):
        p.grad = None
    W.WAN_loss(um, vm, x, 2, 2.0)[0].backward()
torch.cuda.synchronize()
print("ok", pb.ops.last_kernel_path())
