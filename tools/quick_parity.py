"""Development check (GPU): configs 2 and 3 on the tensor-core path (fp32) against the generic fp64 kernel on the same
points and weights: loss and per-tensor gradient errors, relative to the largest gradient entry.
    [PDE_B200_LIB=variant.so] python tools/quick_parity.py [log2_points]"""
import copy
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pde_b200 as pb

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 17
for d, method, bc in ((3, "pinn", "FBC"), (5, "drm", "RB")):
    torch.manual_seed(d)
    n = (1 << lg) + 37
    m = pb.poisson.SolutionNet(d, 64, 5, bc).cuda()
    X = torch.rand(n, d, device="cuda") * 2
    f = pb.poisson.rhs_f_for_u_sin(X, 2.0, [1] * d)
    fn = pb.poisson.pinn_residual_loss if method == "pinn" else pb.poisson.drm_energy_loss
    out = []
    for dt in (torch.float32, torch.float64):
        mm = copy.deepcopy(m).to(dt)
        loss = fn(mm, X.to(dt), f.to(dt), 2.0)
        loss.backward()
        out.append((float(loss), [p.grad.double().clone() for p in mm.parameters()], pb.ops.last_kernel_path()))
    (l32, g32, path), (l64, g64, _) = out
    scale = max(float(g.abs().max()) for g in g64)
    errs = [float((a - b).abs().max()) / scale for a, b in zip(g32, g64)]
    print(f"d={d} {method} n={n} path={path}: loss rel err {abs(l32 - l64) / abs(l64):.2e}, grad err (max-abs / max|g|) {max(errs):.2e}  "
          + " ".join(f"{e:.1e}" for e in errs))
    assert max(errs) < 1e-5 and abs(l32 - l64) < 1e-5 * abs(l64), "parity"
print("quick parity ok")
