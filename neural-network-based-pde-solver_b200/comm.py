"""One-kernel all-reduce over NVLink peer memory for the loss step's exchange (pde_allreduce_oneshot).

    ar = pb.comm.NvlinkAllReduce(group, max_elems=nparam + 8, dtype=torch.float32)
    ar.all_reduce_(buf)          # in place, on the current stream, CUDA-graph capturable

Set-up (once): every rank allocates a peer-visible buffer through the library, the 64-byte CUDA IPC
handles travel through ``torch.distributed.all_gather_object`` and every rank maps its peers' buffers.
The data path never touches NCCL; ``torch.distributed`` is plumbing for the handle exchange only.
Single node only (the ranks of one NVSwitch box), world <= 8.

Lockstep: every rank must issue the same sequence of ``all_reduce_`` calls.  A rank waits for its peers for at most
``timeout_s`` (default 600 s, ``PDE_B200_EXCHANGE_TIMEOUT_S`` overrides, 0 = for ever); if that expires its result is
NaN and an error word is set on the device, which ``check()`` turns into an exception — call it wherever the host
synchronises anyway (FusedTrainer and bench.py do).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib as L
from .ops import _stream


class NvlinkAllReduce:
    def __init__(self, group=None, max_elems=1 << 16, dtype=torch.float32, device=None, timeout_s=None):
        lib = L.load()
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > L.MAX_PEERS:
            raise NotImplementedError("one box: at most 8 ranks")
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.dtype = dtype
        self.dt = L.F64 if dtype == torch.float64 else L.F32
        self.slot = (int(max_elems) + 3) // 4 * 4          # slots stay 16-byte aligned
        nbytes = C.c_size_t(0)
        L.check(lib.pde_peer_bytes(self.dt, self.slot, C.byref(nbytes)), "pde_peer_bytes")
        self._own = C.c_void_p(0)
        handle = C.create_string_buffer(64)
        with torch.cuda.device(self.dev):
            L.check(lib.pde_peer_alloc(nbytes.value, C.byref(self._own), handle), "pde_peer_alloc")
        handles = [None] * self.world
        dist.all_gather_object(handles, handle.raw, group=group)
        self.peers = L.Peers()
        self.peers.rank, self.peers.world = self.rank, self.world
        self._opened = []
        with torch.cuda.device(self.dev):
            for r in range(self.world):
                if r == self.rank:
                    self.peers.base[r] = self._own.value
                else:
                    p = C.c_void_p(0)
                    L.check(lib.pde_peer_open(handles[r], C.byref(p)), "pde_peer_open")
                    self.peers.base[r] = p.value
                    self._opened.append(p.value)
        self.seq = torch.zeros(1, dtype=torch.int32, device=self.dev)
        import os
        timeout_s = float(os.environ.get("PDE_B200_EXCHANGE_TIMEOUT_S", "600") if timeout_s is None else timeout_s)
        L.check(lib.pde_set_exchange_timeout(timeout_s), "pde_set_exchange_timeout")
        dist.barrier(group=group)      # every rank has mapped every buffer before the first call

    def all_reduce_(self, t: torch.Tensor):
        if t.dtype != self.dtype or not t.is_contiguous() or t.device != self.dev or t.numel() > self.slot:
            raise ValueError("tensor does not match the communicator (dtype / device / capacity / contiguity)")
        with torch.cuda.device(self.dev):
            L.check(L.load().pde_allreduce_oneshot(C.byref(self.peers), self.dt, t.data_ptr(), t.numel(), self.slot,
                                                   self.seq.data_ptr(), _stream(self.dev)), "pde_allreduce_oneshot")
        return t

    def check(self):
        """Synchronise the current stream and raise if any exchange of this rank ever timed out (its result was
        poisoned with NaN; the peers' replicas have diverged from this one)."""
        n = C.c_uint32(0)
        with torch.cuda.device(self.dev):
            L.check(L.load().pde_exchange_errors(C.byref(self.peers), C.byref(n), _stream(self.dev)), "pde_exchange_errors")
        if n.value:
            raise L.PdeError(f"NVLink exchange: rank {self.rank} gave up waiting for a peer on {n.value} element(s); "
                             "its gradients were poisoned with NaN and the replicas have diverged")

    def close(self):
        """Unmap the peers' buffers and free the own one (after a barrier: nobody may still read it)."""
        if self._own is None:
            return
        torch.cuda.synchronize(self.dev)
        dist.barrier(group=self.group)
        lib = L.load()
        for p in self._opened:
            lib.pde_peer_close(p)
        dist.barrier(group=self.group)
        lib.pde_peer_free(self._own)
        self._own, self._opened = None, []
