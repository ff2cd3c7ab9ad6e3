"""Latency of the loss step's exchange (51 KB all-reduce of [grad | dE | sums]) — NCCL vs the one-kernel
NVLink all-reduce.  Launch under torchrun: python -m torch.distributed.run --nproc-per-node N tools/exchange_latency.py"""
import json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import pde_b200 as pb

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
world, rank = dist.get_world_size(), dist.get_rank()
n = 12803
buf = torch.randn(n, device="cuda")
ar = pb.comm.NvlinkAllReduce(None, 1 << 15, torch.float32)


def timed(fn, reps=50, inner=20):
    for _ in range(20):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(inner):
            fn()
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / inner * 1e3)
    t = torch.tensor([statistics.median(ts)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def graphed(fn):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            fn()
    return lambda: g.replay()


res = {"world": world, "elements": n, "bytes": 4 * n}
res["nccl_us"] = timed(lambda: dist.all_reduce(buf))
print("nccl done", res, file=sys.stderr, flush=True)
res["nvlink_oneshot_us"] = timed(lambda: ar.all_reduce_(buf))
print("oneshot done", res, file=sys.stderr, flush=True)
buf.normal_()
g1 = graphed(lambda: ar.all_reduce_(buf))
res["nvlink_oneshot_graphed_us"] = timed(g1, reps=30, inner=1) / 20
# the whole data-parallel loss step at a latency-bound size (8192 points, config-2 network): exchange folded into the
# reduction launch vs a separate one-kernel all-reduce vs NCCL
from pde_b200 import ops
torch.manual_seed(0)
m = pb.poisson.SolutionNet(3, 64, 5, "FBC").cuda()
Xs = torch.rand(8192, 3, device="cuda") * 2
fs = pb.poisson.rhs_f_for_u_sin(Xs, 2.0, [1, 1, 1])


def loss_step():
    for p in m.parameters():
        p.grad = None
    pb.poisson.pinn_residual_loss(m, Xs, fs, 2.0, group=dist.group.WORLD).backward()


res["step_nccl_us"] = timed(loss_step, reps=20, inner=10)
ops._EXCHANGE[(dist.group.WORLD, torch.float32)] = ar
ops.FUSE_EXCHANGE = False
res["step_separate_oneshot_us"] = timed(loss_step, reps=20, inner=10)
ops.FUSE_EXCHANGE = True
res["step_fused_exchange_us"] = timed(loss_step, reps=20, inner=10)
ar.check()
ops._EXCHANGE.clear()
if rank == 0:
    print(json.dumps(res))
ar.close()
dist.barrier()
dist.destroy_process_group()
