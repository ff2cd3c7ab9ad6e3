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
if rank == 0:
    print(json.dumps(res))
ar.close()
dist.barrier()
dist.destroy_process_group()
