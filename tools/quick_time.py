"""Quick device timing of the fused loss step (development aid; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pde_b200 as pb

def run(d, method, bc, N, dtype=torch.float32, iters=5):
    torch.manual_seed(0)
    L = 2.0
    m = pb.poisson.SolutionNet(d, 64, 5, bc).to("cuda", dtype)
    X = torch.rand(N, d, device="cuda", dtype=dtype) * L
    f = pb.poisson.rhs_f_for_u_sin(X, L, [1] * d)
    fn = pb.poisson.pinn_residual_loss if method == "pinn" else pb.poisson.drm_energy_loss
    for _ in range(2):
        m.zero_grad(); fn(m, X, f, L).backward()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        m.zero_grad(); fn(m, X, f, L).backward()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"d={d} {method} {bc} N={N} {dtype}: {ms:.3f} ms/step  {N / ms * 1e3:.3e} pts/s", flush=True)

if __name__ == "__main__":
    import os
    if os.environ.get("PDE_B200_LIB"):      # variant builds hold configs 2 and 3 only
        run(3, "pinn", "FBC", 1 << 22, iters=5)
        run(5, "drm", "RB", 1 << 20, iters=10)
        sys.exit(0)
    run(3, "pinn", "FBC", 1 << 18)
    run(3, "pinn", "FBC", 1 << 20)
    run(3, "pinn", "FBC", 1 << 22, iters=3)
    run(5, "drm", "RB", 1 << 20)
    run(5, "pinn", "FBC", 1 << 20)          # 7 jet channels: dimension-split tensor-core passes
    run(1, "pinn", "FBC", 20000)
    run(3, "pinn", "FBC", 1 << 16, dtype=torch.float64, iters=2)
