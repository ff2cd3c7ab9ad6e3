"""Multi-GPU tests (need >= 2 GPUs on the box: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`;
skipped on a single-GPU box).  One process per GPU, rendezvous on 127.0.0.1; torch.distributed (NCCL) is the
plumbing and the comparison, the data path under test is the one-kernel NVLink all-reduce and the
data-parallel fused epoch."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(fn, world, *args):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_entry, args=(world, port, fn, args), nprocs=world, join=True)


def _entry(rank, world, port, fn, args):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        fn(rank, world, *args)
    finally:
        dist.barrier()
        dist.destroy_process_group()


def _need(n):
    if torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")


# ---------------------------------------------------------------- workers (module level: picklable)
def _w_allreduce(rank, world):
    import torch.distributed as dist
    import pde_b200 as pb
    for dtype in (torch.float32, torch.float64):
        ar = pb.comm.NvlinkAllReduce(None, 20000, dtype)
        g = torch.Generator(device="cuda").manual_seed(100 + rank)
        for it, n in enumerate((12803, 1, 20000, 777, 12803, 12803)):
            x = torch.randn(n, dtype=dtype, device="cuda", generator=g) * (10.0 ** (it % 3))
            parts = [torch.empty_like(x) for _ in range(world)]
            dist.all_gather(parts, x)
            want = parts[0].clone()
            for p in parts[1:]:
                want += p                       # rank order 0..W-1, like the kernel
            y = x.clone()
            ar.all_reduce_(y)
            assert torch.equal(y, want), (dtype, n, float((y - want).abs().max()))
            z = x.clone()
            dist.all_reduce(z)
            tol = 1e-5 if dtype == torch.float32 else 1e-13
            assert float((y - z).abs().max()) <= tol * float(z.abs().max())
        # CUDA-graph replay: the call counter lives on the device
        x = torch.full((4096,), float(rank + 1), dtype=dtype, device="cuda")
        ar.all_reduce_(x)                       # warm-up outside the graph
        torch.cuda.synchronize()
        src = torch.full((4096,), float(rank + 1), dtype=dtype, device="cuda")
        g1 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g1):
            x.copy_(src)
            ar.all_reduce_(x)
        for _ in range(5):
            g1.replay()
        torch.cuda.synchronize()
        assert torch.equal(x, torch.full_like(x, world * (world + 1) / 2))
        ar.close()


def _w_trainer(rank, world, exchange):
    import pde_b200 as pb
    L, ks, n_loc, epochs = 2.0, [1, 1], 3000, 5
    torch.manual_seed(21)
    ref = pb.poisson.SolutionNet(2, 32, 4, "FBC").double().cuda()
    X = torch.rand(n_loc * world, 2, dtype=torch.float64, device="cuda") * L      # same on every rank (same seed)
    f = pb.poisson.rhs_f_for_u_sin(X, L, ks)
    import copy
    m = copy.deepcopy(ref)
    tr = pb.train.FusedTrainer(m, L, ks, method="PINN", X=X[rank * n_loc:(rank + 1) * n_loc], f=f[rank * n_loc:(rank + 1) * n_loc],
                               history=epochs, group=torch.distributed.group.WORLD, exchange=exchange)
    tr.step(epochs)
    single = pb.train.FusedTrainer(ref, L, ks, method="PINN", X=X, f=f, history=epochs)
    single.step(epochs)
    np.testing.assert_allclose(tr.hist_loss.cpu().numpy(), single.hist_loss.cpu().numpy(), rtol=1e-11)
    for a, b in zip(m.parameters(), ref.parameters()):
        assert float((a - b).abs().max()) <= 1e-11 * max(1.0, float(b.abs().max()))
    # replicas stay bit-identical
    flat = torch.cat([p.detach().reshape(-1) for p in m.parameters()])
    parts = [torch.empty_like(flat) for _ in range(world)]
    torch.distributed.all_gather(parts, flat)
    assert all(torch.equal(parts[0], p) for p in parts[1:])


def _w_loss_api(rank, world):
    """pinn_residual_loss(..., group=) with the exchange routed through the NVLink kernel equals the single big batch."""
    import pde_b200 as pb
    L, n_loc = 2.0, 5000
    torch.manual_seed(3)
    m = pb.poisson.SolutionNet(3, 64, 5, "FBC").cuda()
    X = torch.rand(n_loc * world, 3, device="cuda") * L
    f = pb.poisson.rhs_f_for_u_sin(X, L, [1, 1, 1])
    big = pb.poisson.pinn_residual_loss(m, X, f, L); big.backward()
    want = [p.grad.clone() for p in m.parameters()]
    m.zero_grad()
    pb.ops.use_nvlink_exchange(None, 1 << 15, torch.float32)
    sl = slice(rank * n_loc, (rank + 1) * n_loc)
    part = pb.poisson.pinn_residual_loss(m, X[sl], f[sl], L, group=torch.distributed.group.WORLD, n_global=n_loc * world)
    part.backward()
    assert abs(float(part) - float(big)) <= 2e-6 * abs(float(big))
    for p, w in zip(m.parameters(), want):
        assert float((p.grad - w).abs().max()) <= 1e-5 * float(w.abs().max())


def _w_loss_api_5d(rank, world):
    """5-D PINN (two dimension-split tensor-core passes): the exchange rides the second reduction only."""
    import pde_b200 as pb
    L, n_loc = 2.0, 4096
    torch.manual_seed(5)
    m = pb.poisson.SolutionNet(5, 64, 4, "FBC").cuda()
    X = torch.rand(n_loc * world, 5, device="cuda") * L
    f = pb.poisson.rhs_f_for_u_sin(X, L, [1] * 5)
    big = pb.poisson.pinn_residual_loss(m, X, f, L); big.backward()
    assert pb.ops.last_kernel_path() == "tcgen05"
    want = [p.grad.clone() for p in m.parameters()]
    m.zero_grad()
    pb.ops.use_nvlink_exchange(None, 1 << 15, torch.float32)
    sl = slice(rank * n_loc, (rank + 1) * n_loc)
    part = pb.poisson.pinn_residual_loss(m, X[sl], f[sl], L, group=torch.distributed.group.WORLD)   # n_global defaults to n x world
    part.backward()
    assert pb.ops.last_kernel_path() == "tcgen05"
    assert abs(float(part) - float(big)) <= 2e-6 * abs(float(big))
    for p, w in zip(m.parameters(), want):
        assert float((p.grad - w).abs().max()) <= 1e-5 * float(w.abs().max())
    pb.ops._EXCHANGE[(torch.distributed.group.WORLD, torch.float32)].check()


def test_nvlink_allreduce_matches_rank_ordered_sum():
    _need(2)
    _run(_w_allreduce, min(torch.cuda.device_count(), 8))


@pytest.mark.parametrize("exchange", ["nvlink", "nccl"])
def test_data_parallel_fused_epoch_equals_single_big_batch(exchange):
    _need(2)
    _run(_w_trainer, 2, exchange)


def test_loss_api_with_nvlink_exchange():
    _need(2)
    _run(_w_loss_api, 2)


def test_loss_api_5d_split_with_nvlink_exchange():
    _need(2)
    _run(_w_loss_api_5d, 2)
