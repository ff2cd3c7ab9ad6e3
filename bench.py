#!/usr/bin/env python
"""bench.py — collocation points / second per train step (fwd + Laplacian + parameter gradient).

Workload (BASELINE.json configs[1]): Poisson_ND.py 3-D PINN, SolutionNet(3, 64, 5, 'FBC'),
2^22 uniform collocation points per GPU, fp32, synthetic points, random-init weights.
A "step" is what the reference does between ``opt.zero_grad()`` and ``opt.step()``:
``loss = pinn_residual_loss(model, X, f, L); loss.backward()`` (+ the gradient all-reduce when
world > 1).  Point sampling, rhs evaluation and Adam are outside the step.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference            # the reference's nested-autograd algorithm on host cores
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

DIM, WIDTH, DEPTH, L_DOM = 3, 64, 5, 2.0
N_PER_GPU = 1 << 22
WORKLOAD = "Poisson_ND 3-D PINN FBC, SolutionNet(3,64,5) sin, 2^22 pts/GPU, fp32"
METRIC = "collocation points/sec per train step (fwd+Δu+param-grad)"
# algorithmic FLOPs per point: 3 (fwd,dgrad,wgrad) * 3 hidden GEMMs * 2*64*64 * 7 channels + small layers
FLOP_PER_POINT = 3 * 3 * 2 * WIDTH * WIDTH * (1 + 2 * DIM) + 6 * WIDTH * (1 + 2 * DIM) + 4 * DIM * WIDTH  # 519 552
CPU_CHUNK = 1 << 16


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except Exception:
                continue
            for nme, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference_rate(n_chunks, threads=None):
    """The reference's algorithm (nested autograd, oracle/autograd_ref.py) on the host cores."""
    from oracle import autograd_ref as AR
    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(0)
    net = AR.build_mlp([DIM] + [WIDTH] * (DEPTH - 1) + [1], "sin", torch.float32)
    X = torch.rand(n_chunks * CPU_CHUNK, DIM) * L_DOM
    f = AR.manufactured_rhs(X, L_DOM, [1] * DIM)
    AR.loss_and_grads("pinn", net, X[:CPU_CHUNK], f[:CPU_CHUNK], L_DOM, "FBC", chunk=CPU_CHUNK)  # warm-up
    t0 = time.perf_counter()
    AR.loss_and_grads("pinn", net, X, f, L_DOM, "FBC", chunk=CPU_CHUNK)
    dt = time.perf_counter() - t0
    return X.shape[0] / dt, dt


def gpu_eager_reference_rate(dev, n_chunks=4):
    """The same nested-autograd algorithm run by PyTorch eager on the B200 itself (SURVEY.md §8d's second,
    recommended baseline): what switching the reference to `device='cuda'` gives without this library."""
    from oracle import autograd_ref as AR
    torch.manual_seed(0)
    net = AR.build_mlp([DIM] + [WIDTH] * (DEPTH - 1) + [1], "sin", torch.float32).to(dev)
    X = torch.rand(n_chunks * CPU_CHUNK, DIM, device=dev) * L_DOM
    f = AR.manufactured_rhs(X, L_DOM, [1] * DIM)
    AR.loss_and_grads("pinn", net, X[:CPU_CHUNK], f[:CPU_CHUNK], L_DOM, "FBC", chunk=CPU_CHUNK)  # warm-up
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    AR.loss_and_grads("pinn", net, X, f, L_DOM, "FBC", chunk=CPU_CHUNK)
    e1.record()
    torch.cuda.synchronize(dev)
    return X.shape[0] / (e0.elapsed_time(e1) * 1e-3)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import autograd_ref as AR
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    net = AR.build_mlp([DIM] + [WIDTH] * (DEPTH - 1) + [1], "sin", torch.float32)
    n = CPU_CHUNK  # bounded sample of the 2^22-point workload per step
    X = torch.rand(n, DIM) * L_DOM
    f = AR.manufactured_rhs(X, L_DOM, [1] * DIM)
    for _ in range(args.warmup):
        AR.loss_and_grads("pinn", net, X, f, L_DOM, "FBC")
    t0 = time.perf_counter()
    for _ in range(args.steps):
        AR.loss_and_grads("pinn", net, X, f, L_DOM, "FBC")
    dt = time.perf_counter() - t0
    val = n * args.steps / dt
    sample = f"{n} of the 2^22 points per step (one 2^16 chunk), oracle/autograd_ref.py nested-autograd port, fp32"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "points/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": val, "unit": "points/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--points", type=int, default=N_PER_GPU, help="points per GPU (default 2^22, the metric's config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    import pde_b200 as pb
    from pde_b200 import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the collocation kernels")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    exchange = "none (one rank)"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries the one JSON line only
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
        # exchange step: one kernel over NVLink peer memory (pde_allreduce_oneshot); PDE_B200_EXCHANGE=nccl keeps NCCL
        if os.environ.get("PDE_B200_EXCHANGE", "nvlink") == "nvlink":
            try:
                ops.use_nvlink_exchange(None, 1 << 15, torch.float32)
                exchange = "one-kernel NVLink peer-memory all-reduce (pde_allreduce_oneshot)"
            except Exception as exc:             # e.g. CUDA IPC unavailable in this container
                exchange = f"NCCL all-reduce (NVLink exchange unavailable: {type(exc).__name__})"
        else:
            exchange = "NCCL all-reduce"
    N = args.points
    n_global = N * world

    torch.manual_seed(0)                       # identical weights on every rank
    model = pb.poisson.SolutionNet(DIM, WIDTH, DEPTH, "FBC").to(dev)
    params = list(model.parameters())
    # 4 rotating point sets (4 x 64 MiB = 256 MiB > 126 MB L2) so that no step finds its inputs in L2
    POOL = 4
    gen = torch.Generator(device=dev); gen.manual_seed(1234 + rank)
    Xs = [torch.rand(N, DIM, device=dev, generator=gen) * L_DOM for _ in range(POOL)]
    fs = [pb.poisson.rhs_f_for_u_sin(X, L_DOM, [1] * DIM) for X in Xs]

    def step(i):
        for p in params:
            p.grad = None
        loss = pb.poisson.pinn_residual_loss(model, Xs[i % POOL], fs[i % POOL], L_DOM, group=group, n_global=n_global)
        loss.backward()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()

    # ---- timed region: device-resident inputs
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ops.KERNEL_EVENTS = []                     # (start, stop) CUDA events around each fused-kernel ABI call
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    kev = ops.KERNEL_EVENTS
    ops.KERNEL_EVENTS = None
    kernel_ms = [a.elapsed_time(b) for a, b in kev]
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = n_global * args.steps / (ms * 1e-3)

    # ---- end to end: host (pinned) inputs each step, H2D on a copy stream one step ahead, loss read back
    hX = [X.cpu().pin_memory() for X in Xs[:2]]
    hf = [f.cpu().pin_memory() for f in fs[:2]]
    dX = [torch.empty_like(Xs[0]) for _ in range(2)]
    df = [torch.empty_like(fs[0]) for _ in range(2)]
    copy_stream = torch.cuda.Stream(dev)
    ready = [torch.cuda.Event() for _ in range(2)]
    freed = [torch.cuda.Event() for _ in range(2)]
    host_loss = torch.empty(1, dtype=torch.float32).pin_memory()

    def upload(i):
        s = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[s])
            dX[s].copy_(hX[s], non_blocking=True)
            df[s].copy_(hf[s], non_blocking=True)
            ready[s].record(copy_stream)

    def e2e_steps(k):
        for s in range(2):
            freed[s].record(torch.cuda.current_stream())
        upload(0)
        for i in range(k):
            s = i % 2
            if i + 1 < k:
                upload(i + 1)
            torch.cuda.current_stream().wait_event(ready[s])
            for p in params:
                p.grad = None
            loss = pb.poisson.pinn_residual_loss(model, dX[s], df[s], L_DOM, group=group, n_global=n_global)
            loss.backward()
            freed[s].record(torch.cuda.current_stream())
            host_loss.copy_(loss.detach().reshape(1), non_blocking=True)
            torch.cuda.current_stream().synchronize()   # the caller reads the loss every step

    e2e_steps(2)
    barrier()
    t0 = time.perf_counter()
    e2e_steps(args.steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = n_global * args.steps / float(t.item())
    h2d = (hX[0].numel() + hf[0].numel()) * 4

    if rank == 0:
        peaks, how = measured_peaks()
        k_ms = statistics.mean(kernel_ms) if kernel_ms else ms / args.steps
        achieved = FLOP_PER_POINT * N / (k_ms * 1e-3) / 1e12
        peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
        prof = {}
        try:
            with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as fh:
                prof = json.load(fh)
        except Exception:
            pass
        out = {
            "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "points_per_gpu": N, "global_points": n_global,
                       "parallelism": f"dp{world} over points, one all-reduce of [grad|dE|sum] (51 KB) per step", "exchange": exchange,
                       "l2": "4 rotating point sets (256 MiB) > 126 MB L2", "kernel_path": pb.ops.last_kernel_path()},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "points/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "how": "pinned host X,f -> device on a copy stream one step ahead; loss .item()-style readback every step"},
            "gpu_launches": 3 * args.steps,   # per step: tc_pack_kernel, tc_kernel<3,2,sin>, reduce_kernel
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": prof.get("dram_bytes_per_launch"),
                         "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({how})",
                         "kernel_ms": k_ms, "flop_per_point": FLOP_PER_POINT,
                         "tensor_pipe_active_pct_ncu": prof.get("tensor_pipe_active_pct"),
                         "traffic_source": prof.get("source"),
                         "hbm_achieved_gbs": (4 * (DIM + 1) * N) / (k_ms * 1e-3) / 1e9},
        }
        if world == 1 and not args.no_cpu_baseline:
            rate, dt = cpu_reference_rate(n_chunks=64, threads=os.cpu_count())
            out["cpu_baseline"] = {"value": rate, "unit": "points/s", "cores": torch.get_num_threads(), "kind": "port",
                                   "sample": f"{64 * CPU_CHUNK} of the 2^22 points (the whole step) in 2^16 chunks ({dt:.1f} s), "
                                             "oracle/autograd_ref.py (the reference's nested-autograd algorithm), fp32"}
            try:
                out["torch_eager_gpu_baseline"] = {
                    "value": gpu_eager_reference_rate(dev), "unit": "points/s",
                    "sample": f"{4 * CPU_CHUNK} points in 2^16 chunks, the same nested-autograd algorithm run by PyTorch eager on this GPU, fp32"}
            except Exception as exc:      # reported extra, never fatal
                out["torch_eager_gpu_baseline"] = {"value": None, "error": type(exc).__name__}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
