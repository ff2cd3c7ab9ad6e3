"""CPU, world_size 2, gloo: the host-side data-parallel logic (pde_b200.ops.combine_forward /
combine_backward).  Each rank stands in for its GPU with the numpy oracle on its own shard of the
points and the combination must reproduce the single-big-batch loss and gradient exactly, both for
plain-mean losses (PINN) and for functions of means (Rayleigh quotient), where averaging per-rank
gradients would be wrong."""
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import flat
from oracle import jets_numpy as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _net(rng, d=2, w=12, depth=4):
    Ws = [rng.uniform(-1, 1, (w, d))] + [rng.uniform(-1, 1, (w, w)) / math.sqrt(w) for _ in range(depth - 2)] + [rng.uniform(-1, 1, (1, w))]
    bs = [rng.uniform(-0.5, 0.5, W.shape[0]) for W in Ws]
    return Ws, bs


def _problem():
    rng = np.random.default_rng(11)
    Ws, bs = _net(rng)
    N = 301                      # uneven shards: 151 + 150
    X = rng.uniform(0.1, 1.9, (N, 2))
    f = rng.normal(size=(N, 1))
    beta = rng.uniform(0.5, 1.5, (N, 1))
    return Ws, bs, X, f, beta


ENV = {"kind": O.ENV_POLY, "lo": 0.0, "hi": 2.0}


def _rank_sums_and_G(Ws, bs, X, programs, order):
    """What one launch of the fused kernel returns for a shard: raw sums of q_k and the raw
    gradient vectors G_k = sum_p dq_k/dtheta (flat, parameters() order)."""
    sums, Gs = [], []
    for prog in programs:
        def program(U, prog=prog):
            q, Ubar = prog(U)
            return float(q.sum()), Ubar, None
        s, gW, gb, _ = O._loss_and_grads(Ws, bs, X, O.SIN, order, ENV, program)
        sums.append(s); Gs.append(flat(gW, gb))
    return np.array(sums), np.stack(Gs)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import pde_b200 as pb
    from pde_b200 import ops
    Ws, bs, X, f, beta = _problem()
    N = X.shape[0]
    sl = slice(0, 151) if rank == 0 else slice(151, N)
    Xr, fr, br = X[sl], f[sl], beta[sl]
    d = X.shape[1]
    res = {}

    # ---- plain mean (PINN): fused launch returns [G/N | dE | sum] -> one all-reduce
    pinn = [lambda U: O.pinn_program(U, d, fr, alpha=-1.0)[:2]]
    sums, G = _rank_sums_and_G(Ws, bs, Xr, pinn, 2)
    nparam = G.shape[1]
    buf = torch.tensor(np.concatenate([G[0] / N, [0.0], sums]))
    means = ops.combine_forward(buf, nparam, float(N), dist.group.WORLD, fused=True)
    res["pinn_loss"] = float(means[0]); res["pinn_grad"] = buf[:nparam].numpy().copy()

    # ---- function of means (Rayleigh quotient m1/m2): sums first, then seeded reverse sweep
    ray = [lambda U: (O.rayleigh_program(U, d, 0.5, br)[0][0], O.rayleigh_program(U, d, 0.5, br)[1][0]),
           lambda U: (O.rayleigh_program(U, d, 0.5, br)[0][1], O.rayleigh_program(U, d, 0.5, br)[1][1])]
    sums, G = _rank_sums_and_G(Ws, bs, Xr, ray, 1)
    buf = torch.tensor(np.concatenate([np.zeros(nparam + 1), sums]))
    means = ops.combine_forward(buf, nparam, float(N), dist.group.WORLD, fused=False)
    m1, m2 = float(means[0]), float(means[1])
    seed = np.array([1.0 / m2, -m1 / (m2 * m2)])             # dF/dm_k, identical on every rank
    buf2 = torch.tensor(np.concatenate([(seed[:, None] * G).sum(0) / N, [0.0], sums]))
    g, _ = ops.combine_backward(buf2, nparam, dist.group.WORLD)
    res["ray_loss"] = m1 / m2; res["ray_grad"] = g.numpy().copy()
    # what naive DDP-style averaging of per-rank loss gradients would give (must differ)
    n_r = Xr.shape[0]
    local_F_grad = ((np.array([1.0 / (sums[1] / n_r), -(sums[0] / n_r) / (sums[1] / n_r) ** 2])[:, None]) * G).sum(0) / n_r
    t = torch.tensor(local_F_grad * n_r / N)
    dist.all_reduce(t)
    res["ray_naive"] = t.numpy().copy()
    # host-side defaults of the data-parallel API: the global point count when the caller gives none, and the
    # sampler's per-rank Philox counter block
    assert ops._global_count(151, dist.group.WORLD, None) == 151 * world
    assert ops._global_count(151, dist.group.WORLD, 301) == 301 and ops._global_count(151, None, None) is None
    from pde_b200.train import rank_offset
    assert rank_offset(0) == 0 and rank_offset(rank) == rank << 40 and rank_offset(1) > 10 ** 9
    if rank == 0:
        np.savez(out, **res)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_combination_is_exact(tmp_path):
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    Ws, bs, X, f, beta = _problem()
    N, d = X.shape
    want, gW, gb, _ = O.eigen_pinn_loss(Ws, bs, X, O.SIN, ENV, -1.0, None, 0.0, f=f)
    assert abs(got["pinn_loss"] - want) <= 1e-12 * abs(want)
    assert np.max(np.abs(got["pinn_grad"] - flat(gW, gb))) <= 1e-12 * np.max(np.abs(flat(gW, gb)))
    want, gW, gb, _ = O.rayleigh_loss(Ws, bs, X, O.SIN, ENV, 0.5, beta)
    ref = flat(gW, gb)
    assert abs(got["ray_loss"] - want) <= 1e-12 * abs(want)
    assert np.max(np.abs(got["ray_grad"] - ref)) <= 1e-11 * np.max(np.abs(ref))
    # the shortcut the combination avoids really is wrong
    assert np.max(np.abs(got["ray_naive"] - ref)) > 1e-6 * np.max(np.abs(ref))
