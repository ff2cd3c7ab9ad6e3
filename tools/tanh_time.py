"""Development aid (GPU): loss-step time of a tanh network on the tcgen05 path (2-D PINN, [2,64,64,64,64,1], 2^20 points)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pde_b200 as pb
from pde_b200 import _lib as L
from pde_b200.ops import ProgramSpec, residual_means
from pde_b200.schrodinger._common import poly_envelope

torch.manual_seed(0)
mods = []
for i, o in ((2, 64), (64, 64), (64, 64), (64, 64)):
    mods += [torch.nn.Linear(i, o), torch.nn.Tanh()]
net = torch.nn.Sequential(*mods, torch.nn.Linear(64, 1)).cuda()
m = torch.nn.Module(); m.net = net
N = 1 << 20
X = torch.rand(N, 2, device="cuda") * 2.0
spec = ProgramSpec(L.PROG_PINN, alpha=1.0, beta_const=4.9)
def step():
    for p in net.parameters(): p.grad = None
    residual_means(m, X, spec, poly_envelope(2.0))[0].backward()
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): step()
e1.record(); torch.cuda.synchronize()
print(f"tanh 2-D PINN N=2^20 ({pb.ops.last_kernel_path()}): {e0.elapsed_time(e1) / 10:.3f} ms/step")
