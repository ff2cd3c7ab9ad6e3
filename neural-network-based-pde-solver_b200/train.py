"""Fused training epoch around the loss step.

``train_poisson_nd`` (Poisson_ND.py:215-300) spends an epoch on: [draw points, evaluate f] -> loss ->
``loss.backward()`` -> ``Adam.step()`` -> draw test points -> L2 error -> ``.item()`` -> keep the best
state on the CPU.  Here the same epoch is a fixed sequence of launches on one stream —

    pde_sample_points_rhs  (optional, when points are redrawn every epoch)
    pde_residual_loss_grad (the fused tcgen05 / SIMT loss step: sums, flat gradient)
    all-reduce of [grad | dE | sums]  (only with ``group``)
    pde_adam_step          (fused Adam on the flat gradient, parameters updated in place)
    pde_sample_points_rhs + pde_residual_loss_grad(value only) + pde_keep_best  (optional evaluation)

— captured once into a CUDA graph and replayed; nothing is read back to the host unless the caller
asks for ``loss`` / ``l2``.  Sampling is statistically, not bitwise, equivalent to ``torch.rand``
(Philox4x32-10 keyed by seed, counter = point index and epoch); everything else follows the
reference's arithmetic (Adam: torch.optim.Adam defaults, single-tensor formula).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib as L
from .ops import NO_ENVELOPE, EnvelopeSpec, ProgramSpec, _Net, _stream, _ws_for


def sample_points_rhs(n, dim, L_box, ks=None, *, dtype=torch.float32, device="cuda", seed=0, offset=0, lo=0.0,
                      want_u=False, want_f=True, X=None):
    """(X, u_exact, f): uniform points in [lo, L_box)^dim (or the given ``X``) with the manufactured
    solution / right-hand side of Poisson_ND.py:49-58 evaluated in the same launch."""
    lib = L.load()
    dev = torch.device(device)
    if X is not None:
        Xin = X.detach().contiguous()
        n, dim, dtype, dev = Xin.shape[0], Xin.shape[1], Xin.dtype, Xin.device
        Xout = Xin
    else:
        Xin = None
        Xout = torch.empty(n, dim, dtype=dtype, device=dev)
    u = torch.empty(n, 1, dtype=dtype, device=dev) if (want_u and ks is not None) else None
    f = torch.empty(n, 1, dtype=dtype, device=dev) if (want_f and ks is not None) else None
    karr = (C.c_double * dim)(*[float(k) for k in ks]) if ks is not None else None
    with torch.cuda.device(dev):
        L.check(lib.pde_sample_points_rhs(L.F64 if dtype == torch.float64 else L.F32, dim, n, float(lo), float(L_box),
                                          int(seed), int(offset), None, karr, float(L_box),
                                          Xin.data_ptr() if Xin is not None else None,
                                          None if Xin is not None else Xout.data_ptr(),
                                          u.data_ptr() if u is not None else None,
                                          f.data_ptr() if f is not None else None, _stream(dev)),
                "pde_sample_points_rhs")
    return Xout, u, f


class FusedAdam:
    """Adam over a list of parameter tensors driven by one flat gradient vector (pde_adam_step).
    State tensors mirror torch.optim.Adam's ``exp_avg`` / ``exp_avg_sq`` / ``step``."""

    def __init__(self, params: Sequence[torch.Tensor], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.params = list(params)
        if not self.params or len(self.params) > 2 * L.MAX_LINEAR + 1:
            raise NotImplementedError("1 .. 17 parameter tensors")
        p0 = self.params[0]
        for p in self.params:
            if not p.is_cuda or p.dtype != p0.dtype or p.device != p0.device or not p.is_contiguous():
                raise ValueError("parameters must be contiguous CUDA tensors of one dtype on one device")
        self.n = sum(p.numel() for p in self.params)
        self.exp_avg = torch.zeros(self.n, dtype=p0.dtype, device=p0.device)
        self.exp_avg_sq = torch.zeros(self.n, dtype=p0.dtype, device=p0.device)
        self.step_count = torch.zeros(1, dtype=torch.int64, device=p0.device)
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay

    def config(self, grad_scale=1.0) -> L.Adam:
        c = L.Adam()
        c.dtype = L.F64 if self.params[0].dtype == torch.float64 else L.F32
        c.n_tensors = len(self.params)
        c.lr, c.beta1, c.beta2, c.eps = float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps)
        c.weight_decay, c.grad_scale = float(self.weight_decay), float(grad_scale)
        for i, p in enumerate(self.params):
            c.param[i] = p.data_ptr()
            c.numel[i] = p.numel()
        return c

    def step(self, grad_flat: torch.Tensor, grad_scale=1.0):
        """``grad_flat``: at least ``self.n`` contiguous values in parameters() order."""
        dev = self.params[0].device
        if grad_flat.numel() < self.n or grad_flat.dtype != self.params[0].dtype or not grad_flat.is_contiguous():
            raise ValueError("flat gradient has the wrong size / dtype")
        cfg = self.config(grad_scale)
        with torch.cuda.device(dev):
            L.check(L.load().pde_adam_step(C.byref(cfg), grad_flat.data_ptr(), self.exp_avg.data_ptr(),
                                           self.exp_avg_sq.data_ptr(), self.step_count.data_ptr(), _stream(dev)),
                    "pde_adam_step")


class FusedTrainer:
    """The PINN / DRM branch of ``train_poisson_nd`` (Poisson_ND.py:215-241,281-300) with every epoch
    replayed from one CUDA graph.

    model      SolutionNet-style module on a CUDA device (parameters updated in place)
    method     'PINN' | 'DRM'
    X, f       fixed interior points and right-hand side; default: drawn once like the reference
               (:193-194), or every epoch with ``resample=True`` (what the WAN branch does, :246,:256)
    n_test     > 0: L2 evaluation on freshly drawn test points after each step and device-side
               best-parameter tracking (:281-300)
    history    number of epochs whose loss (and L2) are kept in device arrays (0: none)
    group      data parallelism over points: every rank runs the same epoch on its own points and the
               [grad | dE | sums] vector is summed over the ranks before Adam (exchange='nvlink': one
               kernel over peer memory; 'nccl': torch.distributed.all_reduce)
    """

    def __init__(self, model, L_box=2.0, ks=None, method='PINN', n_interior=20000, lr=1e-3, betas=(0.9, 0.999), eps=1e-8,
                 weight_pde=1.0, X=None, f=None, resample=False, seed=0, n_test=0, history=0, graph=True, group=None,
                 envelope: Optional[EnvelopeSpec] = None, exchange='nvlink'):
        from .poisson import _envelope
        self.lib = L.load()
        self.model, self.L, self.method, self.group = model, float(L_box), method, group
        p0 = next(model.parameters())
        self.dev, self.dtype = p0.device, p0.dtype
        if method not in ('PINN', 'DRM'):
            raise ValueError("method must be one of {'PINN','DRM'}")
        self.env = envelope if envelope is not None else _envelope(model, L_box)
        self.spec = ProgramSpec(L.PROG_PINN, alpha=-1.0) if method == 'PINN' else ProgramSpec(L.PROG_DRM, alpha=0.5)
        self.dim = getattr(model, "dim", None) or _Net(model, torch.empty(1, 1, device=self.dev, dtype=self.dtype)).dim
        self.ks = [1.0] * self.dim if ks is None else [float(k) for k in ks]
        self.seed, self.resample = int(seed), bool(resample)
        if X is None:
            self.n = int(n_interior)
            self.X, _, self.f = sample_points_rhs(self.n, self.dim, self.L, self.ks, dtype=self.dtype, device=self.dev,
                                                  seed=self.seed, offset=0)
        else:
            self.X = X.detach().to(self.dtype).contiguous()
            self.n = self.X.shape[0]
            self.f = (f if f is not None else sample_points_rhs(0, 0, self.L, self.ks, X=self.X)[2]).detach().to(self.dtype).reshape(-1).contiguous()
            if self.resample:
                raise ValueError("resample=True draws its own points; do not pass X")
        self.net = _Net(model, self.X)
        self.params = [p for p in self.net.params]
        self.nparam = sum(p.numel() for p in self.params)
        self.opt = FusedAdam([p.data for p in self.params], lr=lr, betas=betas, eps=eps)
        self.weight_pde = float(weight_pde)
        self.world = 1
        self._ar = None
        if group is not None:
            import torch.distributed as dist
            self.world = dist.get_world_size(group)
            if exchange == 'nvlink':     # one-kernel all-reduce over peer memory (pde_allreduce_oneshot)
                from .comm import NvlinkAllReduce
                self._ar = NvlinkAllReduce(group, self.nparam + 2, self.dtype, self.dev)
            elif exchange != 'nccl':
                raise ValueError("exchange must be 'nvlink' or 'nccl'")
        # [grad (nparam) | dE (1) | sums (K)] — the layout pde_residual_loss_grad writes and Adam reads
        self.buf = torch.zeros(self.nparam + 2, dtype=self.dtype, device=self.dev)
        self.n_test = int(n_test)
        if self.n_test:
            self.Xt = torch.empty(self.n_test, self.dim, dtype=self.dtype, device=self.dev)
            self.ut = torch.empty(self.n_test, dtype=self.dtype, device=self.dev)
            self.esum = torch.zeros(1, dtype=self.dtype, device=self.dev)
            self.best_metric = torch.full((1,), float("inf"), dtype=self.dtype, device=self.dev)
            self.best_flat = torch.zeros(self.nparam, dtype=self.dtype, device=self.dev)
            self.best_step = torch.full((1,), -1, dtype=torch.int64, device=self.dev)
        self.hist_loss = torch.zeros(max(int(history), 0), dtype=self.dtype, device=self.dev)
        self.hist_l2sq = torch.zeros(max(int(history), 0) if self.n_test else 0, dtype=self.dtype, device=self.dev)
        self._karr = (C.c_double * self.dim)(*self.ks)
        self._order = self.lib.pde_program_order(self.spec.kind)
        cnet = self.net.to_c([p.data for p in self.params])
        self._ws = _ws_for(cnet, self._order, max(self.n, self.n_test, 1), self.dev)
        self._graph = None
        self._use_graph = bool(graph)
        self.epochs_done = 0

    # ---- the launches of one epoch (enqueued on the current stream, no host sync)
    def _enqueue(self):
        lib, dev = self.lib, self.dev
        st = _stream(dev)
        ps = [p.data for p in self.params]
        cnet = self.net.to_c(ps)
        cenv = self.env.to_c()
        dt = L.F64 if self.dtype == torch.float64 else L.F32
        es = self.buf.element_size()
        if self.resample:
            L.check(lib.pde_sample_points_rhs(dt, self.dim, self.n, 0.0, self.L, self.seed, 0, self.opt.step_count.data_ptr(),
                                              self._karr, self.L, None, self.X.data_ptr(), None, self.f.data_ptr(), st),
                    "pde_sample_points_rhs")
        prog = L.Program()
        prog.kind, prog.alpha = self.spec.kind, self.spec.alpha
        prog.f = self.f.data_ptr()
        inv_n = 1.0 / (self.n * self.world)
        L.check(lib.pde_residual_loss_grad(C.byref(cnet), C.byref(cenv), C.byref(prog), self.X.data_ptr(), self.n, None, inv_n,
                                           self.buf.data_ptr() + (self.nparam + 1) * es, self.buf.data_ptr(),
                                           self.buf.data_ptr() + self.nparam * es, self._ws.data_ptr(), self._ws.numel(), st),
                "pde_residual_loss_grad")
        if self._ar is not None:
            self._ar.all_reduce_(self.buf)
        elif self.group is not None:
            import torch.distributed as dist
            dist.all_reduce(self.buf, group=self.group)
        if self.hist_loss.numel():
            self.hist_loss.index_copy_(0, self.opt.step_count.clamp(max=self.hist_loss.numel() - 1), self.buf[self.nparam + 1:] * inv_n)
        cfg = self.opt.config(self.weight_pde)
        L.check(lib.pde_adam_step(C.byref(cfg), self.buf.data_ptr(), self.opt.exp_avg.data_ptr(), self.opt.exp_avg_sq.data_ptr(),
                                  self.opt.step_count.data_ptr(), st), "pde_adam_step")
        if self.n_test:
            # fresh test points every epoch (Poisson_ND.py:282), a different Philox key than the interior draw
            L.check(lib.pde_sample_points_rhs(dt, self.dim, self.n_test, 0.0, self.L, self.seed ^ 0x9E3779B97F4A7C15, 0,
                                              self.opt.step_count.data_ptr(), self._karr, self.L, None, self.Xt.data_ptr(),
                                              self.ut.data_ptr(), None, st), "pde_sample_points_rhs")
            ev = L.Program()
            ev.kind, ev.alpha = L.PROG_MSE, 1.0
            ev.f = self.ut.data_ptr()
            L.check(lib.pde_residual_loss_grad(C.byref(cnet), C.byref(cenv), C.byref(ev), self.Xt.data_ptr(), self.n_test, None,
                                               1.0 / self.n_test, self.esum.data_ptr(), None, None, self._ws.data_ptr(),
                                               self._ws.numel(), st), "pde_residual_loss_grad(eval)")
            L.check(lib.pde_keep_best(C.byref(cfg), self.esum.data_ptr(), self.best_metric.data_ptr(), self.best_flat.data_ptr(),
                                      self.opt.step_count.data_ptr(), self.best_step.data_ptr(), st), "pde_keep_best")
            if self.hist_l2sq.numel():
                self.hist_l2sq.index_copy_(0, (self.opt.step_count - 1).clamp(min=0, max=self.hist_l2sq.numel() - 1),
                                           self.esum / self.n_test)

    def step(self, n_epochs=1):
        """Run ``n_epochs`` epochs (graph replays after the first call)."""
        with torch.cuda.device(self.dev):
            for _ in range(int(n_epochs)):
                if not self._use_graph:
                    self._enqueue()
                elif self._graph is None:
                    # warm-up epoch outside the graph (module loading, cudaFuncSetAttribute), then capture
                    self._enqueue()
                    torch.cuda.synchronize(self.dev)
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        self._enqueue()
                    self._graph = g
                    self.epochs_done += 1
                    continue
                else:
                    self._graph.replay()
                self.epochs_done += 1
        return self

    # ---- read-backs (these synchronise)
    @property
    def loss(self):
        """PDE loss of the last evaluated epoch (mean over the global batch), 0-d device tensor."""
        return (self.buf[self.nparam + 1] / (self.n * self.world)).clone()

    @property
    def l2(self):
        """L2 error on the last test draw: sqrt(mean((u - u*)^2))   (Poisson_ND.py:285)."""
        if not self.n_test:
            raise ValueError("constructed with n_test=0")
        return (self.esum[0] / self.n_test).sqrt()

    @property
    def best_l2(self):
        return (self.best_metric[0] / self.n_test).sqrt()

    def load_best(self, model=None):
        """Copy the best parameters seen so far into ``model`` (default: the trained model)."""
        tgt = self.params if model is None else _Net(model, self.X).params
        o = 0
        with torch.no_grad():
            for p in tgt:
                k = p.numel()
                p.copy_(self.best_flat[o:o + k].view_as(p)); o += k


class GraphedEpoch:
    """Capture a reference-style epoch — ``zero_grad -> losses -> backward -> optimizer.step`` written
    against the drop-in loss functions — into one CUDA graph and replay it.

    The drop-in operators enqueue everything on the current stream and never synchronise, so the host
    side of an epoch (Python, autograd bookkeeping, ~20 small launches per loss) disappears on replay;
    this is what makes the latency-bound configurations (1 000 – 40 000 points: IPW / QHO / KH grids,
    BASELINE.json configs 1, 4, 5) run at kernel speed.  Requirements on ``fn``: static input tensors,
    no ``.item()`` / host read-back inside, optimisers built with ``capturable=True``, gradients
    allocated before capture (``warmup`` eager epochs on a side stream take care of that).

        opt = torch.optim.Adam(model.parameters(), lr=1e-3, capturable=True)
        def epoch():
            opt.zero_grad(set_to_none=False)
            loss = I.PINN_loss(model, x, n, L); loss.backward(); opt.step()
            return loss
        ep = pb.train.GraphedEpoch(epoch); ep(); ...; print(float(ep.out))
    """

    def __init__(self, fn, warmup=3, device=None):
        self.fn = fn
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.graph = None
        self.out = None
        self.warmup = int(warmup)
        self.replays = 0

    def _capture(self):
        side = torch.cuda.Stream(self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            for _ in range(self.warmup):
                self.out = self.fn()
        torch.cuda.current_stream(self.dev).wait_stream(side)
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.out = self.fn()
        self.graph = g

    def __call__(self):
        """Run one epoch; returns ``fn``'s output tensors (static storage, overwritten by every replay).
        The first call runs ``warmup`` eager epochs and captures; it does not replay, so that the number
        of optimiser steps taken equals ``warmup`` after it (then +1 per call)."""
        with torch.cuda.device(self.dev):
            if self.graph is None:
                self._capture()
            else:
                self.graph.replay()
                self.replays += 1
        return self.out
