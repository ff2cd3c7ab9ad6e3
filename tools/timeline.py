"""Development aid (GPU): per-step timeline of CTA 0 (epilogue warp 0 and the MMA issuer).
Needs a library built with `make -C neural-network-based-pde-solver_b200/csrc EXTRA=-DPDE_TC_TIMELINE`.
    PDE_B200_TIMELINE=gpurun_out/timeline.txt python tools/timeline.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
out = os.environ.setdefault("PDE_B200_TIMELINE", "gpurun_out/timeline.txt")
import torch, pde_b200 as pb
torch.manual_seed(0)
cfg3 = len(sys.argv) > 1 and sys.argv[1] == "cfg3"       # 5-D Deep Ritz (6 channels, W streamed) instead of config 2
d = 5 if cfg3 else 3
m = pb.poisson.SolutionNet(d, 64, 5, "RB" if cfg3 else "FBC").cuda()
N = 1 << 20
X = torch.rand(N, d, device="cuda") * 2
f = pb.poisson.rhs_f_for_u_sin(X, 2.0, [1] * d)
for _ in range(2):
    m.zero_grad()
    l = (pb.poisson.drm_energy_loss if cfg3 else pb.poisson.pinn_residual_loss)(m, X, f, 2.0); l.backward()
torch.cuda.synchronize()
ev = [[], []]
wev = {}
for line in open(out):
    w, i, c = line.split()
    if int(w) >= 10:
        wev.setdefault(int(w) - 10, []).append((int(i), int(c)))
    else:
        ev[int(w)].append((int(i), int(c)))
names = {1: "tile start", 2: "program done", 3: "  after barrier 1", 4: "  X in smem, after barrier 2", 5: "  x^T written, loads issued"}
for k in range(5):
    names[10 + k] = f"F{k} enter"; names[20 + k] = f"F{k} D ready"; names[30 + k] = f"B{k} enter"; names[40 + k] = f"B{k} Ab ready"
    names[50 + k] = f"B{k} want sets"; names[60 + k] = f"B{k} sets free"
    names[300 + k] = f"  F1 chunk {k} top"; names[310 + k] = f"  F1 chunk {k} tmem ready"; names[320 + k] = f"  F1 chunk {k} math done"
    names[330 + k] = f"  F1 chunk {k} stsm issued"; names[340 + k] = f"  F1 chunk {k} fenced+arrived"
e = ev[0]
starts = [i for i, (id_, _) in enumerate(e) if id_ == 1]
if len(starts) > 6:
    a, b = starts[4], starts[5]
    t0 = e[a][1]
    print("epilogue warp 0 (cycles since tile start; delta)")
    prev = t0
    for id_, c in e[a:b + 1]:
        print(f"  {names.get(id_, id_):28s} {c - t0:8d}  +{c - prev}")
        prev = c
    print("issuer (100+10l+j: fwd chunk j seen, 200+10l+j: dgrad chunk j seen)")
    for id_, c in ev[1]:
        if t0 <= c <= e[b][1]:
            print(f"  {id_:4d} {c - t0:8d}")

if wev:
    # arrival of every epilogue warp at the CTA-wide barriers, relative to the first warp to arrive
    print("per-warp arrival at the barriers (cycles after the first warp; id 1: tile start, 2: end of forward, 3: residual stage, 4: after it)")
    nt = min(len([1 for i, _ in wev[w] if i == 1]) for w in wev)
    for t in range(3, min(nt, 9)):
        for bid in (1, 2, 3, 4):
            arr = []
            for w in sorted(wev):
                xs = [c for i, c in wev[w] if i == bid]
                if t < len(xs):
                    arr.append(xs[t])
            if arr:
                print(f"  tile {t} barrier {bid}: " + " ".join(f"{c - min(arr):6d}" for c in arr))
