"""CPU emulation of the tensor-core operand splits (development aid): rounds GEMM operands the way
the tcgen05 path does (x = hi + lo in bf16 / fp16, fp32 accumulate) inside the forward / dgrad /
wgrad contractions of the 3-D PINN step and reports loss / gradient error vs float64."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import jets_numpy as O

def split(x, fmt, terms):
    t = torch.from_numpy(x.astype(np.float32))
    dt = torch.bfloat16 if fmt == "bf16" else torch.float16
    hi = t.to(dt).float()
    if terms == 1:
        return [hi.numpy().astype(np.float64)]
    lo = (t - hi).to(dt).float()
    return [hi.numpy().astype(np.float64), lo.numpy().astype(np.float64)]

def mm(a, b, fmt, mode):
    """a @ b with split operands. mode: 'fp32' exact, '1', '3' (hh+hl+lh), '4'."""
    if mode == "fp32":
        return (a.astype(np.float32) @ b.astype(np.float32)).astype(np.float64)
    A = split(a, fmt, 1 if mode == "1" else 2); B = split(b, fmt, 1 if mode == "1" else 2)
    out = A[0] @ B[0]
    if mode in ("3", "4"):
        out = out + A[0] @ B[1] + A[1] @ B[0]
    if mode == "4":
        out = out + A[1] @ B[1]
    return out.astype(np.float32).astype(np.float64)

def run(fmt, mode, N=16384, d=3, seed=0, bscale=1.0):
    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    import pde_b200  # noqa
    from pde_b200.poisson import SolutionNet, rhs_f_for_u_sin
    m = SolutionNet(d, 64, 5, "FBC")
    lin = [x for x in m.net if isinstance(x, torch.nn.Linear)]
    Ws = [l.weight.detach().double().numpy() for l in lin]; bs = [l.bias.detach().double().numpy() for l in lin]
    X = rng.uniform(0, 2, (N, d)); f = rhs_f_for_u_sin(torch.tensor(X), 2.0, [1] * d).numpy()
    ref_loss, rW, rb = O.poisson_pinn_loss(Ws, bs, X, f, 2.0, "FBC")
    # emulated: same algorithm, hidden GEMMs through mm()
    C = 1 + 2 * d; n = len(Ws)
    a = [X] + [np.broadcast_to(np.eye(d)[i], (N, d)) for i in range(d)] + [np.zeros((N, d)) for _ in range(d)]
    Zs = []; As = []
    for l in range(n - 1):
        W, b = Ws[l], bs[l]
        if l == 0:
            z = [ai @ W.T for ai in a]
        else:
            z = [mm(ai, W.T, fmt, mode) for ai in a]
        z[0] = z[0] + b
        Zs.append(z); As.append(a)
        s0, s1, s2, s3 = O._act(z[0], O.SIN)
        a = [s0] + [s1 * z[1 + i] for i in range(d)] + [s2 * z[1 + i] ** 2 + s1 * z[1 + d + i] for i in range(d)]
        a = [x.astype(np.float32).astype(np.float64) for x in a]
    J = np.concatenate([ai @ Ws[-1].T + (bs[-1] if k == 0 else 0) for k, ai in enumerate(a)], axis=1)
    U, ctx = O.apply_envelope(J, X, 2, O.ENV_POLY, 0.0, 2.0)
    q, Ub, _ = O.pinn_program(U, d, f, alpha=-1.0)
    loss = q.mean()
    Jb = O.envelope_backward(Ub / N * bscale, ctx, 2, d)
    gW = [None] * n; gb = [None] * n
    zb = [Jb[:, c:c + 1] for c in range(C)]
    gW[-1] = sum(zb[c].T @ a[c] for c in range(C)); gb[-1] = zb[0].sum(0)
    ab = [zb[c] @ Ws[-1] for c in range(C)]
    for l in range(n - 2, -1, -1):
        z = Zs[l]
        s0, s1, s2, s3 = O._act(z[0], O.SIN)
        nz = [None] * C
        nz[0] = s1 * ab[0]
        for i in range(d):
            nz[0] = nz[0] + s2 * z[1 + i] * ab[1 + i] + (s3 * z[1 + i] ** 2 + s2 * z[1 + d + i]) * ab[1 + d + i]
            nz[1 + i] = s1 * ab[1 + i] + 2 * s2 * z[1 + i] * ab[1 + d + i]
            nz[1 + d + i] = s1 * ab[1 + d + i]
        nz = [x.astype(np.float32).astype(np.float64) for x in nz]
        ain = As[l]
        if l == 0:
            gW[0] = sum(nz[c].T @ ain[c] for c in range(C))
        else:
            gW[l] = sum(mm(nz[c].T, ain[c], fmt, mode) for c in range(C))
            ab = [mm(nz[c], Ws[l], fmt, mode) for c in range(C)]
        gb[l] = nz[0].sum(0)
    gW = [g / bscale for g in gW]; gb = [g / bscale for g in gb]
    gref = np.concatenate([np.concatenate([w.ravel(), b.ravel()]) for w, b in zip(rW, rb)])
    gem = np.concatenate([np.concatenate([w.ravel(), b.ravel()]) for w, b in zip(gW, gb)])
    per = max(np.linalg.norm(a - b) / np.linalg.norm(b) for a, b in zip(gW + gb, rW + rb) if np.linalg.norm(b) > 0)
    print(f"{fmt} terms={mode}: loss rel {abs(loss - ref_loss) / abs(ref_loss):.2e}  grad rel-l2 {np.linalg.norm(gem - gref) / np.linalg.norm(gref):.2e}"
          f"  worst tensor {per:.2e}  max-abs/max {np.max(np.abs(gem - gref)) / np.max(np.abs(gref)):.2e}")

if __name__ == "__main__":
    for fmt, mode, sc in [("bf16", "3", 1.0), ("fp16", "3", 1.0), ("fp16", "3", 16384.0), ("fp16", "3", 16384.0 * 64)]:
        print("bscale", sc, end=": "); run(fmt, mode, bscale=sc)
