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
                      want_u=False, want_f=True, X=None, offset_add=None, out=None):
    """(X, u_exact, f): uniform points in [lo, L_box)^dim (or the given ``X``) with the manufactured
    solution / right-hand side of Poisson_ND.py:49-58 evaluated in the same launch.  ``offset`` selects the
    Philox counter block: draws with different offsets are independent (epoch number; ``rank_offset(r)`` for the
    ranks of a data-parallel run).  ``offset_add``: device int64 scalar added to ``offset`` at run time (a draw
    counter, so that a replayed CUDA graph draws fresh points); ``out=(X, u, f)``: write into existing tensors."""
    lib = L.load()
    dev = torch.device(device)
    if X is not None:
        Xin = X.detach().contiguous()
        n, dim, dtype, dev = Xin.shape[0], Xin.shape[1], Xin.dtype, Xin.device
        Xout = Xin
    else:
        Xin = None
        Xout = out[0] if out is not None else torch.empty(n, dim, dtype=dtype, device=dev)
    if out is not None:
        u, f = out[1], out[2]
    else:
        u = torch.empty(n, 1, dtype=dtype, device=dev) if (want_u and ks is not None) else None
        f = torch.empty(n, 1, dtype=dtype, device=dev) if (want_f and ks is not None) else None
    karr = (C.c_double * dim)(*[float(k) for k in ks]) if ks is not None else None
    with torch.cuda.device(dev):
        L.check(lib.pde_sample_points_rhs(L.F64 if dtype == torch.float64 else L.F32, dim, n, float(lo), float(L_box),
                                          int(seed), int(offset), offset_add.data_ptr() if offset_add is not None else None,
                                          karr, float(L_box),
                                          Xin.data_ptr() if Xin is not None else None,
                                          None if Xin is not None else Xout.data_ptr(),
                                          u.data_ptr() if u is not None else None,
                                          f.data_ptr() if f is not None else None, _stream(dev)),
                "pde_sample_points_rhs")
    return Xout, u, f


def rank_offset(rank):
    """Philox counter offset of data-parallel rank ``rank``: bits 40.. of the counter's second half, far above any
    epoch count, so that every rank draws its own points (the reference is single-process; with the default seed
    identical draws on every rank would make the global batch n points repeated world times)."""
    return int(rank) << 40


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


class Adam:
    """``torch.optim.Adam(params, lr, betas, eps, weight_decay)`` for reference-style epochs (``zero_grad -> loss ->
    backward -> step``) at two or three launches of bookkeeping per step instead of torch's ~45: ``step()`` is one
    ``pde_adam_step`` launch over all parameter tensors, fed by one flat gradient vector (same update rule and state as
    ``FusedAdam``).  Where that vector comes from depends on how the gradients were zeroed:

    * ``zero_grad()`` (``set_to_none=True``, torch's default): no launch.  The drop-in losses hand autograd views of
      *one* flat gradient buffer (``ops._Residual`` / ``ops._Jets``), which autograd adopts as ``p.grad`` without
      copying when ``p.grad`` is None; ``step()`` recognises that layout (same storage, consecutive offsets in
      parameter order) and passes the buffer to the kernel as it is.  A second ``backward()`` accumulates into the
      adopted views in place, which keeps the layout; several operator terms inside one ``backward()`` are summed per
      parameter by the autograd engine before they reach ``p.grad`` and are gathered (below).
    * ``zero_grad(set_to_none=False)``: the ``.grad`` tensors are views of the optimiser's own flat buffer, zeroed by
      one fill; autograd adds into them (one small launch per parameter tensor).

    Gradients that arrive in any other layout (other operators, a different parameter order, hooks) are gathered into
    the optimiser's buffer first (one ``torch.cat``; a parameter without gradient counts as zero gradient — torch.optim
    would skip it).  Nothing synchronises, so an epoch can be captured in a ``GraphedEpoch`` as it is (no ``capturable``
    flag needed); the arithmetic is that of torch's default Adam, bias corrections in double (``capturable=True`` keeps
    its step count in float32 and is ~6e-6 per step away from it)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.params = [p for p in params]
        self._opt = FusedAdam(self.params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        p0 = self.params[0]
        self.flat = torch.zeros(self._opt.n, dtype=p0.dtype, device=p0.device)
        self._off = []
        off = 0
        for p in self.params:
            self._off.append(off)
            off += p.numel()
        self.adopted_steps = 0     # steps that took the gradients where autograd left them (no gather, no fill)
        self.gathered_steps = 0

    @property
    def state(self):
        return {"exp_avg": self._opt.exp_avg, "exp_avg_sq": self._opt.exp_avg_sq, "step": self._opt.step_count}

    def _own_view(self, i):
        p, off = self.params[i], self._off[i]
        return self.flat[off:off + p.numel()].view_as(p)

    def zero_grad(self, set_to_none=True):
        if set_to_none:
            for p in self.params:
                p.grad = None
            return
        es = self.flat.element_size()
        for i, p in enumerate(self.params):
            g = p.grad
            if g is None or g.data_ptr() != self.flat.data_ptr() + self._off[i] * es or not g.is_contiguous():
                p.grad = self._own_view(i)
        self.flat.zero_()

    def _flat_grads(self):
        """The gradients as one flat vector without copying, or None if they are not laid out that way."""
        g0 = self.params[0].grad
        if g0 is None:
            return None
        es = g0.element_size()
        st = g0.untyped_storage()
        base = g0.data_ptr()
        for p, off in zip(self.params, self._off):
            g = p.grad
            if (g is None or g.dtype != g0.dtype or not g.is_contiguous() or g.data_ptr() != base + off * es
                    or g.untyped_storage().data_ptr() != st.data_ptr()):
                return None
        if (g0.storage_offset() + self._opt.n) * es > st.nbytes():
            return None
        return g0.as_strided((self._opt.n,), (1,), g0.storage_offset())

    def step(self):
        flat = self._flat_grads()
        if flat is None:
            zero = None
            parts = []
            for i, p in enumerate(self.params):
                if p.grad is None:
                    if zero is None:
                        zero = torch.zeros(max(q.numel() for q in self.params), dtype=self.flat.dtype, device=self.flat.device)
                    parts.append(zero[:p.numel()])
                else:
                    parts.append(p.grad.reshape(-1))
            flat = torch.cat(parts)      # (not into self.flat: some of the parts may be views of it)
            self.gathered_steps += 1
        else:
            self.adopted_steps += 1
        self._opt.step(flat)


class FusedTrainer:
    """The PINN / DRM branch of ``train_poisson_nd`` (Poisson_ND.py:215-241,281-300) with every epoch
    replayed from one CUDA graph: PDE term plus, when their weights are non-zero, the soft Dirichlet
    penalty of ``bc_mode='RB'`` (:224-228, fresh face points every epoch), the data term (:230-234) and
    the norm term (:236-237); total gradient = sum of weight x term gradient, then Adam.

    model      SolutionNet-style module on a CUDA device (parameters updated in place)
    method     'PINN' | 'DRM'
    X, f       fixed interior points and right-hand side; default: drawn once like the reference
               (:193-194), or every epoch with ``resample=True`` (what the WAN branch does, :246,:256)
    weights    {'pde','bc','data','norm'} like the reference's ``weights`` argument; defaults follow :169-173
               (bc 1e4 for an 'RB' model, data 1e3 when data points are given, norm 0)
    n_boundary face points per epoch, split evenly over the 2d faces (:225)
    X_data, u_data   fixed data points / values (:197-202)
    norm_mode  'nontrivial' (1 / (mean u^2 + 1e-8)) | 'l2' (mean u^2) on the interior points (:143-147)
    n_test     > 0: L2 evaluation on freshly drawn test points after each step and device-side
               best-parameter tracking (:281-300)
    history    number of epochs whose loss (and L2) are kept in device arrays (0: none)
    group      data parallelism over points: every rank runs the same epoch on its own points (the rank selects
               the sampler's Philox counter block, ``rank_offset``) and the
               [grad | dE | sums] rows are summed over the ranks before Adam (exchange='nvlink': one kernel over
               peer memory; 'nccl': torch.distributed.all_reduce)
    """

    TERMS = ('pde', 'bc', 'data', 'norm')

    def __init__(self, model, L_box=2.0, ks=None, method='PINN', n_interior=20000, lr=1e-3, betas=(0.9, 0.999), eps=1e-8,
                 weight_pde=1.0, X=None, f=None, resample=False, seed=0, n_test=0, history=0, graph=True, group=None,
                 envelope: Optional[EnvelopeSpec] = None, exchange='nvlink', weights=None, n_boundary=4000, X_data=None,
                 u_data=None, norm_mode='nontrivial'):
        from .poisson import _envelope
        self.lib = L.load()
        self.model, self.L, self.method, self.group = model, float(L_box), method, group
        p0 = next(model.parameters())
        self.dev, self.dtype = p0.device, p0.dtype
        if method not in ('PINN', 'DRM'):
            raise ValueError("method must be one of {'PINN','DRM'}")
        if norm_mode not in ('nontrivial', 'l2'):
            raise ValueError("norm mode should be 'nontrivial' or 'l2'")
        self.env = envelope if envelope is not None else _envelope(model, L_box)
        self.spec = ProgramSpec(L.PROG_PINN, alpha=-1.0) if method == 'PINN' else ProgramSpec(L.PROG_DRM, alpha=0.5)
        self.dim = getattr(model, "dim", None) or _Net(model, torch.empty(1, 1, device=self.dev, dtype=self.dtype)).dim
        self.ks = [1.0] * self.dim if ks is None else [float(k) for k in ks]
        self.seed, self.resample = int(seed), bool(resample)
        self.world, self.rank = 1, 0
        if group is not None:
            import torch.distributed as dist
            self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if X is None:
            self.n = int(n_interior)
            self.X, _, self.f = sample_points_rhs(self.n, self.dim, self.L, self.ks, dtype=self.dtype, device=self.dev,
                                                  seed=self.seed, offset=rank_offset(self.rank))
        else:
            self.X = X.detach().to(self.dtype).contiguous()
            self.n = self.X.shape[0]
            self.f = (f if f is not None else sample_points_rhs(0, 0, self.L, self.ks, X=self.X)[2]).detach().to(self.dtype).reshape(-1).contiguous()
            if self.resample:
                raise ValueError("resample=True draws its own points; do not pass X")
        # term weights (Poisson_ND.py:169-173)
        rb = getattr(model, "bc_mode", "FBC") == 'RB' and envelope is None
        self.w = {'pde': float(weight_pde), 'bc': 1e4 if rb else 0.0, 'data': 1e3 if X_data is not None else 0.0, 'norm': 0.0}
        if weights:
            unknown = set(weights) - set(self.TERMS)
            if unknown:
                raise ValueError(f"unknown loss terms {sorted(unknown)}")
            self.w.update({k: float(v) for k, v in weights.items()})
        self.norm_mode = norm_mode
        self.use = {'pde': True, 'bc': self.w['bc'] != 0.0, 'data': self.w['data'] != 0.0, 'norm': self.w['norm'] > 0.0}
        if self.use['data'] and (X_data is None or u_data is None):
            raise ValueError("a data weight needs X_data and u_data")
        if self.use['norm'] and norm_mode == 'nontrivial' and group is not None:
            raise NotImplementedError("the 'nontrivial' norm term is a function of a global mean: not available with group "
                                      "(use norm_mode='l2' or weight 0)")
        self.rows = [t for t in self.TERMS if self.use[t]]
        self.net = _Net(model, self.X)
        self.params = [p for p in self.net.params]
        self.nparam = sum(p.numel() for p in self.params)
        self.opt = FusedAdam([p.data for p in self.params], lr=lr, betas=betas, eps=eps)
        self.weight_pde = self.w['pde']
        self._ar = None
        # one row per active term: [grad (nparam) | dE (1) | sum (1)] — the layout pde_residual_loss_grad writes
        self.G = torch.zeros(len(self.rows), self.nparam + 2, dtype=self.dtype, device=self.dev)
        self.buf = self.G[0]
        if group is not None:
            if exchange == 'nvlink':     # one-kernel all-reduce over peer memory (pde_allreduce_oneshot)
                from .comm import NvlinkAllReduce
                self._ar = NvlinkAllReduce(group, self.G.numel(), self.dtype, self.dev)
            elif exchange != 'nccl':
                raise ValueError("exchange must be 'nvlink' or 'nccl'")
        if len(self.rows) > 1:
            self.total = torch.zeros(self.nparam + 2, dtype=self.dtype, device=self.dev)
            self.wvec = torch.tensor([self.w[t] for t in self.rows], dtype=self.dtype, device=self.dev)
        if self.use['bc']:
            per_face = max(1, int(n_boundary) // (2 * self.dim))       # Poisson_ND.py:225
            self.nb = 2 * self.dim * per_face
            self.Xb = torch.empty(self.nb, self.dim, dtype=self.dtype, device=self.dev)
            # face k = (coordinate k // 2, side k % 2): that coordinate is replaced by 0 or L  (:133-139)
            keep = torch.ones(2 * self.dim, 1, self.dim, dtype=self.dtype, device=self.dev)
            put = torch.zeros(2 * self.dim, 1, self.dim, dtype=self.dtype, device=self.dev)
            for k in range(2 * self.dim):
                keep[k, 0, k // 2] = 0.0
                put[k, 0, k // 2] = self.L if k % 2 else 0.0
            self._face_keep = keep.expand(-1, per_face, -1).reshape(self.nb, self.dim).contiguous()
            self._face_put = put.expand(-1, per_face, -1).reshape(self.nb, self.dim).contiguous()
        if self.use['data']:
            self.Xd = X_data.detach().to(self.dtype).to(self.dev).contiguous()
            self.ud = u_data.detach().to(self.dtype).to(self.dev).reshape(-1).contiguous()
        if self.use['norm']:
            self.nsum = torch.zeros(1, dtype=self.dtype, device=self.dev)
            self.nseed = torch.ones(1, dtype=self.dtype, device=self.dev)
            self.norm_value = torch.zeros((), dtype=self.dtype, device=self.dev)
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
        n_big = max(self.n, self.n_test, getattr(self, "nb", 1), self.Xd.shape[0] if self.use['data'] else 1)
        self._ws = _ws_for(cnet, self._order, n_big, self.dev)
        self._graph = None
        self._use_graph = bool(graph)
        self.epochs_done = 0

    # ---- the launches of one epoch (enqueued on the current stream, no host sync)
    def _mse(self, cnet, cenv, X, n, target, inv_n, row, seed=None, sums=None, want_grad=True):
        """mean((u - target)^2) (target None: mean u^2) with gradient into row ``row`` of self.G."""
        lib, st = self.lib, _stream(self.dev)
        es = self.G.element_size()
        base = self.G.data_ptr() + row * self.G.shape[1] * es
        ev = L.Program()
        ev.kind, ev.alpha = L.PROG_MSE, 1.0
        ev.f = target.data_ptr() if target is not None else None
        L.check(lib.pde_residual_loss_grad(C.byref(cnet), C.byref(cenv), C.byref(ev), X.data_ptr(), n,
                                           seed.data_ptr() if seed is not None else None, inv_n,
                                           sums.data_ptr() if sums is not None else base + (self.nparam + 1) * es,
                                           base if want_grad else None, None, self._ws.data_ptr(), self._ws.numel(), st),
                "pde_residual_loss_grad(mse)")

    def _enqueue(self):
        lib, dev = self.lib, self.dev
        st = _stream(dev)
        ps = [p.data for p in self.params]
        cnet = self.net.to_c(ps)
        cenv = self.env.to_c()
        dt = L.F64 if self.dtype == torch.float64 else L.F32
        es = self.G.element_size()
        P = self.nparam
        if self.resample:
            L.check(lib.pde_sample_points_rhs(dt, self.dim, self.n, 0.0, self.L, self.seed, rank_offset(self.rank),
                                              self.opt.step_count.data_ptr(), self._karr, self.L, None, self.X.data_ptr(), None,
                                              self.f.data_ptr(), st), "pde_sample_points_rhs")
        prog = L.Program()
        prog.kind, prog.alpha = self.spec.kind, self.spec.alpha
        prog.f = self.f.data_ptr()
        inv_n = 1.0 / (self.n * self.world)
        row = self.rows.index('pde')
        base = self.G.data_ptr() + row * self.G.shape[1] * es
        fused_exchange = self._ar is not None and len(self.rows) == 1
        if fused_exchange:     # exchange folded into the reduction of the per-CTA partials (one launch fewer)
            L.check(lib.pde_residual_loss_grad_exchange(C.byref(cnet), C.byref(cenv), C.byref(prog), self.X.data_ptr(), self.n, None,
                                                        inv_n, base, self._ws.data_ptr(), self._ws.numel(), C.byref(self._ar.peers),
                                                        self._ar.slot, self._ar.seq.data_ptr(), st),
                    "pde_residual_loss_grad_exchange")
        else:
            L.check(lib.pde_residual_loss_grad(C.byref(cnet), C.byref(cenv), C.byref(prog), self.X.data_ptr(), self.n, None, inv_n,
                                               base + (P + 1) * es, base, base + P * es, self._ws.data_ptr(), self._ws.numel(), st),
                    "pde_residual_loss_grad")
        if self.use['bc']:
            # fresh face points every epoch: one uniform draw, face coordinate replaced by 0 / L; the mean over
            # the faces of the per-face means (equal counts) is the mean over all face points
            L.check(lib.pde_sample_points_rhs(dt, self.dim, self.nb, 0.0, self.L, self.seed ^ 0x5851F42D4C957F2D, rank_offset(self.rank),
                                              self.opt.step_count.data_ptr(), None, self.L, None, self.Xb.data_ptr(), None, None, st),
                    "pde_sample_points_rhs(faces)")
            torch.addcmul(self._face_put, self.Xb, self._face_keep, out=self.Xb)
            self._mse(cnet, cenv, self.Xb, self.nb, None, 1.0 / (self.nb * self.world), self.rows.index('bc'))
        if self.use['data']:
            self._mse(cnet, cenv, self.Xd, self.Xd.shape[0], self.ud, 1.0 / (self.Xd.shape[0] * self.world), self.rows.index('data'))
        if self.use['norm']:
            r = self.rows.index('norm')
            if self.norm_mode == 'l2':
                self._mse(cnet, cenv, self.X, self.n, None, inv_n, r)
            else:
                # F = 1 / (m + 1e-8), m = mean u^2: the mean first, then the reverse sweep seeded with dF/dm
                self._mse(cnet, cenv, self.X, self.n, None, inv_n, r, sums=self.nsum, want_grad=False)
                m = self.nsum * inv_n
                torch.neg(torch.reciprocal((m + 1e-8) ** 2), out=self.nseed)
                self._mse(cnet, cenv, self.X, self.n, None, inv_n, r, seed=self.nseed)
                self.norm_value.copy_(torch.reciprocal(m + 1e-8)[0])
        if fused_exchange:
            pass
        elif self._ar is not None:
            self._ar.all_reduce_(self.G.view(-1))
        elif self.group is not None:
            import torch.distributed as dist
            dist.all_reduce(self.G, group=self.group)
        if self.hist_loss.numel():
            self.hist_loss.index_copy_(0, self.opt.step_count.clamp(max=self.hist_loss.numel() - 1), self.G[row, P + 1:] * inv_n)
        if len(self.rows) > 1:
            torch.mv(self.G.t(), self.wvec, out=self.total)      # sum of weight x term gradient
            gflat, gscale = self.total, 1.0
        else:
            gflat, gscale = self.G[0], self.w['pde']
        cfg = self.opt.config(gscale)
        L.check(lib.pde_adam_step(C.byref(cfg), gflat.data_ptr(), self.opt.exp_avg.data_ptr(), self.opt.exp_avg_sq.data_ptr(),
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
    def check(self):
        """Raise if the NVLink exchange ever timed out on this rank (synchronises the stream)."""
        if self._ar is not None:
            self._ar.check()

    @property
    def loss(self):
        """PDE loss of the last evaluated epoch (mean over the global batch), 0-d device tensor."""
        self.check()
        return (self.G[self.rows.index('pde'), self.nparam + 1] / (self.n * self.world)).clone()

    @property
    def terms(self):
        """{'pde','bc','data','norm','total'} of the last evaluated epoch as 0-d device tensors
        (what train_poisson_nd appends to its history, Poisson_ND.py:288-293)."""
        self.check()
        P = self.nparam
        out = {t: torch.zeros((), dtype=self.dtype, device=self.dev) for t in self.TERMS}
        out['pde'] = self.G[self.rows.index('pde'), P + 1] / (self.n * self.world)
        if self.use['bc']:
            out['bc'] = self.G[self.rows.index('bc'), P + 1] / (self.nb * self.world)
        if self.use['data']:
            out['data'] = self.G[self.rows.index('data'), P + 1] / (self.Xd.shape[0] * self.world)
        if self.use['norm']:
            out['norm'] = (self.G[self.rows.index('norm'), P + 1] / (self.n * self.world)) if self.norm_mode == 'l2' else self.norm_value
        out['total'] = sum(self.w[t] * out[t] for t in self.TERMS)
        return out

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


class WanTrainer:
    """The WAN branch of ``train_poisson_nd`` (Poisson_ND.py:242-276): per epoch ``critic_steps`` critic updates on
    freshly drawn interior points (``loss_v`` of wan_losses, :244-248), then one solution update on another fresh draw
    with ``w_pde loss_pde_u + w_bc bc + w_data data + w_norm norm`` (:251-271).  The losses are the drop-in operators
    of ``pde_b200.poisson`` (both networks' jets and reverse sweeps on the fused kernels), the optimisers are
    ``pde_b200.train.Adam`` (torch.optim.Adam's arithmetic, one launch per step), the points come from the device
    sampler with a draw counter, and the whole epoch is captured into one CUDA graph (``graph=True``).

    ``record_points=True`` (eager mode only) keeps every draw in ``self.drawn`` = [(kind, X, f) ...] so that a
    reference loop can be run on identical points.
    """

    def __init__(self, model, critic, L_box=2.0, ks=None, n_interior=20000, lr=1e-3, critic_steps=3, wan_reg=1.0,
                 weights=None, n_boundary=4000, X_data=None, u_data=None, norm_mode='nontrivial', seed=0, graph=True,
                 record_points=False):
        from . import poisson as P
        self.P = P
        self.model, self.critic, self.L = model, critic, float(L_box)
        p0 = next(model.parameters())
        self.dev, self.dtype = p0.device, p0.dtype
        self.dim = model.dim
        self.ks = [1.0] * self.dim if ks is None else [float(k) for k in ks]
        self.n, self.critic_steps, self.wan_reg = int(n_interior), int(critic_steps), float(wan_reg)
        rb = getattr(model, "bc_mode", "FBC") == 'RB'
        self.w = {'pde': 1.0, 'bc': 1e4 if rb else 0.0, 'data': 1e3 if X_data is not None else 0.0, 'norm': 0.0}
        if weights:
            self.w.update({k: float(v) for k, v in weights.items()})
        self.norm_mode = norm_mode
        self.seed = int(seed)
        self.opt_u = Adam(model.parameters(), lr=lr)      # torch.optim.Adam's update, one launch per step
        self.opt_v = Adam(critic.parameters(), lr=lr)
        self.u_params, self.v_params = list(model.parameters()), list(critic.parameters())
        self.draws = torch.zeros(1, dtype=torch.int64, device=self.dev)
        # static point buffers (one per evaluation of the epoch, so that the autograd graphs of an epoch do not alias)
        mk = lambda n: (torch.empty(n, self.dim, dtype=self.dtype, device=self.dev).requires_grad_(True), None,
                        torch.empty(n, 1, dtype=self.dtype, device=self.dev))
        self.bufs = [mk(self.n) for _ in range(self.critic_steps + 1)]
        if self.w['bc'] != 0.0:
            per_face = max(1, int(n_boundary) // (2 * self.dim))
            self.nb = 2 * self.dim * per_face
            self.Xb = torch.empty(self.nb, self.dim, dtype=self.dtype, device=self.dev)
            keep = torch.ones(2 * self.dim, 1, self.dim, dtype=self.dtype, device=self.dev)
            put = torch.zeros(2 * self.dim, 1, self.dim, dtype=self.dtype, device=self.dev)
            for k in range(2 * self.dim):
                keep[k, 0, k // 2] = 0.0
                put[k, 0, k // 2] = self.L if k % 2 else 0.0
            self._face_keep = keep.expand(-1, per_face, -1).reshape(self.nb, self.dim).contiguous()
            self._face_put = put.expand(-1, per_face, -1).reshape(self.nb, self.dim).contiguous()
        self.Xd = X_data.detach().to(self.dev, self.dtype).contiguous() if X_data is not None else None
        self.ud = u_data.detach().to(self.dev, self.dtype).contiguous() if u_data is not None else None
        self.record = bool(record_points) and not graph
        self.drawn = []
        self._ep = GraphedEpoch(self._epoch, device=self.dev) if graph else None
        self.last = None

    def _draw(self, buf, kind):
        X, _, f = buf
        sample_points_rhs(self.n, self.dim, self.L, self.ks, dtype=self.dtype, device=self.dev, seed=self.seed,
                          offset_add=self.draws, out=(X.detach(), None, f))
        self.draws.add_(1)
        if self.record:
            self.drawn.append((kind, X.detach().clone(), f.clone()))
        return X, f

    def _epoch(self):
        P = self.P
        loss_v = None
        for k in range(self.critic_steps):                                   # Poisson_ND.py:244-248
            X, f = self._draw(self.bufs[k], 'v')
            _, loss_v, _, _ = P.wan_losses(self.model, self.critic, X, f, self.L, v_reg_weight=self.wan_reg)
            self.opt_v.zero_grad()
            loss_v.backward(inputs=self.v_params)
            self.opt_v.step()
        X, f = self._draw(self.bufs[self.critic_steps], 'u')                 # :251-253
        loss_pde_u, _, weak, phi_n = P.wan_losses(self.model, self.critic, X, f, self.L, v_reg_weight=self.wan_reg)
        zero = torch.zeros((), dtype=self.dtype, device=self.dev)
        bc_l = data_l = norm_l = zero
        if self.w['bc'] != 0.0:                                              # :255-259 on device-drawn face points
            sample_points_rhs(self.nb, self.dim, self.L, None, dtype=self.dtype, device=self.dev,
                              seed=self.seed ^ 0x5851F42D4C957F2D, offset_add=self.draws, out=(self.Xb, None, None))
            self.draws.add_(1)
            torch.addcmul(self._face_put, self.Xb, self._face_keep, out=self.Xb)
            if self.record:
                self.drawn.append(('bc', self.Xb.clone(), None))
            bc_l = P.data_loss(self.model, self.Xb, None, self.L)
        if self.w['data'] != 0.0:                                            # :261-265
            data_l = P.data_loss(self.model, self.Xd, self.ud, self.L)
        if self.w['norm'] > 0.0:                                             # :267-268
            norm_l = P.norm_loss(P.solution_jets(self.model, X.detach(), self.L, 0)[0], mode=self.norm_mode)
        loss = self.w['pde'] * loss_pde_u + self.w['bc'] * bc_l + self.w['data'] * data_l + self.w['norm'] * norm_l
        self.opt_u.zero_grad()
        loss.backward(inputs=self.u_params)
        self.opt_u.step()
        return {'total': loss.detach(), 'pde': loss_pde_u.detach(), 'bc': bc_l.detach(), 'data': data_l.detach(),
                'norm': norm_l.detach(), 'wan_loss_v': loss_v.detach() if loss_v is not None else zero,
                'wan_weak': weak, 'wan_phi_norm': phi_n}

    def step(self, n_epochs=1):
        for _ in range(int(n_epochs)):
            self.last = self._ep() if self._ep is not None else self._epoch()
        return self


class GraphedEpoch:
    """Capture a reference-style epoch — ``zero_grad -> losses -> backward -> optimizer.step`` written
    against the drop-in loss functions — into one CUDA graph and replay it.

    The drop-in operators enqueue everything on the current stream and never synchronise, so the host
    side of an epoch (Python, autograd bookkeeping, ~20 small launches per loss) disappears on replay;
    this is what makes the latency-bound configurations (1 000 – 40 000 points: IPW / QHO / KH grids,
    BASELINE.json configs 1, 4, 5) run at kernel speed.  Requirements on ``fn``: static input tensors,
    no ``.item()`` / host read-back inside, optimisers built with ``capturable=True`` (or ``pb.train.Adam``, which
    updates all parameter tensors in one launch), gradients allocated before capture (``warmup`` eager epochs on a
    side stream take care of that).

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
