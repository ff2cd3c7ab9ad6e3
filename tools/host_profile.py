"""Development aid (GPU): where the host time of an eager drop-in loss step goes (cProfile, 300 steps of config 1)."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pde_b200 as pb
m = pb.poisson.SolutionNet(1, 64, 5, "FBC").cuda()
X = torch.rand(20000, 1, device="cuda") * 2
f = pb.poisson.rhs_f_for_u_sin(X, 2.0, [1])
def step():
    for p in m.parameters():
        p.grad = None
    pb.poisson.pinn_residual_loss(m, X, f, 2.0).backward()
for _ in range(20):
    step()
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(300):
    step()
torch.cuda.synchronize(); print("ms per eager step:", (time.perf_counter() - t) / 300 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(300):
    step()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
