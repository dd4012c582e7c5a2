#!/usr/bin/env python
"""bench.py -- hybrid-rollout throughput of the sm_100a path (and of the CPU reference port).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], per GPU): an ensemble of 4096 initial conditions x
64 cells, stencil radius 3, dt = 1e-3, MODEL_CONFIG weights (F=4, H=128, L=4, random
init, seed 0), smooth random-mode synthetic ICs.  One bench "step" = one hybrid time
step of the whole ensemble = ONE launch of the fused kernel; the K timed steps form a
K-step rollout (each step consumes the previous step's state).  Metric:
cell-updates/s = ICs x cells x steps / device time, summed over all GPUs (weak scaling:
every rank owns its own 4096 ICs, no data-path collective).

Timing: per-step CUDA-event pairs on the launching stream, an L2 flush (256 MiB write)
between steps outside the pairs, barrier + synchronize on both sides of the K steps,
max over ranks.  `e2e` runs the same K steps through HybridSolver.step_pinned(), i.e.
pinned host state -> H2D -> fused step -> D2H every step.

`--impl reference` times the CPU restatement of the reference (oracle/, all host
threads) on a bounded sample of the same workload and prints the same JSON line.
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

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "hybrid_rollout_cell_updates_per_sec"
UNIT = "cell-updates/s"
ICS, NX, RADIUS, DT = 4096, 64, 3, 1e-3
FLOP_PER_CELL_EXECUTED = 329_216      # split edge MLP; what the kernel executes
FLOP_PER_CELL_REFERENCE = 394_752     # as the reference computes it
BYTES_PER_CELL = 24                   # read n,u,E + write n',u',E'
WORKLOAD = f"C2 ensemble rollout: {ICS} ICs x {NX} cells per GPU, radius {RADIUS}, dt={DT:g}, H=128 L=4"


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def make_ics(n_ics: int, first_seed: int) -> np.ndarray:
    from oracle import ref_port as P          # input generator only (the reference's IC family, stabilised)
    grid = P.Grid(nx=NX, dt=DT)
    base = np.stack([P.stable_initial_condition(grid, first_seed + s) for s in range(min(n_ics, 256))])
    reps = (n_ics + len(base) - 1) // len(base)
    ics = np.tile(base, (reps, 1, 1))[:n_ics].copy()
    # make every IC distinct: a small per-IC velocity offset (E stays consistent with n)
    off = np.random.RandomState(first_seed).uniform(-1e-2, 1e-2, size=(n_ics, 1)).astype(np.float32)
    ics[:, 1] += off
    return ics


class ClockSampler:
    """nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _loop(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self._stop.wait(0.15)

    def __enter__(self):
        self._t = threading.Thread(target=self._loop, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# ----------------------------------------------------------------------------- CPU reference port
def cpu_reference_rate(n_ics: int, steps: int, warmup: int):
    """cell-updates/s of the batched CPU port (all host threads) on n_ics ICs."""
    from oracle import batched, ref_port as P
    torch.set_num_threads(os.cpu_count() or 1)
    w = P.init_weights(0)
    grid = P.Grid(nx=NX, dt=DT)
    state = torch.from_numpy(make_ics(n_ics, 0))
    for _ in range(warmup):
        state = batched.hybrid_step(w, state, grid.x, grid.k, grid.dt, grid.dx, radius=RADIUS)
    t0 = time.perf_counter()
    for _ in range(steps):
        state = batched.hybrid_step(w, state, grid.x, grid.k, grid.dt, grid.dx, radius=RADIUS)
    dt = time.perf_counter() - t0
    assert torch.isfinite(state).all()
    return n_ics * NX * steps / dt, dt


def cpu_unbatched_rate(n_ics: int, steps: int):
    """cell-updates/s of the faithful one-IC-at-a-time port (the reference's own loop structure)."""
    from oracle import ref_port as P
    torch.set_num_threads(os.cpu_count() or 1)
    w = P.init_weights(0)
    grid = P.Grid(nx=NX, dt=DT)
    ics = make_ics(n_ics, 0)
    t0 = time.perf_counter()
    for ic in ics:
        s = ic
        for _ in range(steps):
            s = P.hybrid_step(w, s, grid, radius=RADIUS)
    return n_ics * NX * steps / (time.perf_counter() - t0)


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return 0
    sample_ics = 512
    rate, secs = cpu_reference_rate(sample_ics, args.steps, max(args.warmup, 1))
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"{sample_ics} of {ICS} ICs per step"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"batched CPU port (oracle/batched.py, torch {torch.__version__}, {cores} threads), "
                                   f"{sample_ics} ICs x {NX} cells x {args.steps} steps"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------- sm_100a path
def run_ours(args):
    rank, world, local = dist_env()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the sm_100a path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from gnn_plasma_flux_b200 import FluxGNN, HybridSolver, MODEL_CONFIG, _lib
    from oracle import ref_port as P          # weights initialiser + IC generator (inputs only)

    weights = P.init_weights(0)
    model = FluxGNN(**MODEL_CONFIG)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()})
    solver = HybridSolver(None, RADIUS, nx=NX, dt=DT, device=dev, graph_radius=RADIUS, model=model.to(dev),
                          precision=args.precision)
    ics = make_ics(ICS, 1000 * rank)
    K, W = args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    a = torch.from_numpy(ics).to(dev)
    b = torch.empty_like(a)
    stream = torch.cuda.current_stream(dev)

    # ---- device-resident K-step rollout, one launch per step --------------------------
    for _ in range(W):
        solver.rollout(a, 1, out=b)
        a, b = b, a
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    barrier()
    launches0 = _lib.launch_count()
    with ClockSampler(local) as clocks:
        wall0 = time.perf_counter()
        for i in range(K):
            flush.zero_()                                   # L2 flush, outside the timed pair
            starts[i].record(stream)
            solver.rollout(a, 1, out=b)
            stops[i].record(stream)
            a, b = b, a
        barrier()
        wall = time.perf_counter() - wall0
    launches = _lib.launch_count() - launches0
    dev_ms = sum(s.elapsed_time(e) for s, e in zip(starts, stops))
    assert torch.isfinite(a).all(), "rollout went non-finite"
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    value = world * ICS * NX * K / (dev_ms * 1e-3)

    # ---- end to end: pinned host state in, pinned host state out, every step ---------------
    h_in = torch.from_numpy(ics).pin_memory()
    h_out = torch.empty_like(h_in).pin_memory()
    for _ in range(max(W, 1)):
        solver.step_pinned(h_in, h_out)
        h_in, h_out = h_out, h_in
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        solver.step_pinned(h_in, h_out)
        h_in, h_out = h_out, h_in
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * ICS * NX * K / float(t.item())
    state_bytes = ICS * 3 * NX * 4
    # same loop with the kernel reading / writing the pinned host state directly (zero-copy over PCIe)
    for _ in range(max(W, 1)):
        solver.step_pinned(h_in, h_out, zero_copy=True)
        h_in, h_out = h_out, h_in
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        solver.step_pinned(h_in, h_out, zero_copy=True)
        h_in, h_out = h_out, h_in
    barrier()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_zero_copy = world * ICS * NX * K / float(t.item())
    assert torch.isfinite(h_in).all()

    # ---- the tensor-core variants of the same step, reported beside the headline ----------------
    tensor_extra = {}
    if not args.no_tensor_path:
        for prec in ("tf32x3", "tf32"):
            if prec == args.precision:
                continue
            tsol = HybridSolver(None, RADIUS, nx=NX, dt=DT, device=dev, graph_radius=RADIUS, model=solver.model,
                                precision=prec)
            ta = torch.from_numpy(ics).to(dev)
            tb = torch.empty_like(ta)
            for _ in range(W):
                tsol.rollout(ta, 1, out=tb)
                ta, tb = tb, ta
            t_ms = 0.0
            barrier()
            for i in range(K):
                flush.zero_()
                starts[i].record(stream)
                tsol.rollout(ta, 1, out=tb)
                stops[i].record(stream)
                ta, tb = tb, ta
            barrier()
            t_ms = sum(s_.elapsed_time(e_) for s_, e_ in zip(starts, stops))
            tt = torch.tensor([t_ms], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            tensor_extra[prec] = float(tt.item())

    line = None
    if rank == 0:
        # ---- FP32-pipe roofline of this device, measured live ------------------------------
        sink = torch.empty(148 * 8 * 256, dtype=torch.float32, device=dev)
        probe = {0: 0.0, 1: 0.0}
        for packed in (0, 1, 0, 1, 0, 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            flops = _lib.lib().fluxgnn_ffma_probe(sink.data_ptr(), 148 * 8, 256, 20000, packed, stream.cuda_stream)
            e1.record(stream)
            torch.cuda.synchronize(dev)
            probe[packed] = max(probe[packed], flops / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        best = max(probe.values())
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        per_gpu = ICS * NX * K / (dev_ms * 1e-3)
        achieved_tf = per_gpu * FLOP_PER_CELL_EXECUTED / 1e12
        bf16_peak = float(peaks.get("bf16_tflops", 1590.0))
        tensor_path = {}
        for prec, t_ms in tensor_extra.items():
            rate = world * ICS * NX * K / (t_ms * 1e-3)
            products = 3 if prec == "tf32x3" else 1
            executed = rate / world * 327_680 * products / 1e12      # tensor-core FLOPs actually issued per GPU
            tensor_path[prec] = {
                "value": rate, "unit": UNIT, "ms_per_step": t_ms / K, "kernel": "hybrid_tc_kernel<3>",
                "parity": ("same gates as fp32 (<=1e-5 per step, 1000-step gate; flux error vs fp64 ~4e-7)"
                           if prec == "tf32x3" else "looser: flux error ~3e-4, state <=1e-5 per step, <=2e-3 over 1000 steps"),
                "roofline": {"bound": "tensor", "achieved": executed, "peak": bf16_peak / 2, "unit": "TFLOP/s",
                             "frac": executed / (bf16_peak / 2), "traffic": None,
                             "note": f"kind::tf32, {products} UMMA product(s) per contraction; peak = measured bf16 "
                                     "cuBLAS burst / 2 (no direct TF32 measurement in MEASURED_PEAKS.json)"}}
        headline_kernel = "hybrid_tile_kernel<3>" if args.precision == "fp32" else "hybrid_tc_kernel<3>"
        roofline = {
            "bound": "fp32-ffma", "kernel": headline_kernel,
            "achieved": achieved_tf, "peak": best, "unit": "TFLOP/s", "frac": achieved_tf / best if best else None,
            "peak_source": "register-only FMA probe kernels timed in this run (fluxgnn_ffma_probe: FFMA %.1f, FFMA2 %.1f TFLOP/s); "
                           "nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.5" % (probe[0], probe[1]),
            "flop_per_cell": FLOP_PER_CELL_EXECUTED,
            "achieved_reference_flop_count": per_gpu * FLOP_PER_CELL_REFERENCE / 1e12,
            "traffic": None,
            "hbm": {"achieved": per_gpu * BYTES_PER_CELL / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": per_gpu * BYTES_PER_CELL / 1e9 / hbm_peak,
                    "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650",
                    "note": "the hybrid step is FP32-compute bound (16 kFLOP/B); HBM is not the binding roof"},
            "avg_launch_ms": dev_ms / K,
        }
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            _, probe_s = cpu_reference_rate(512, 2, 1)                  # size the sample to ~12 s of CPU work
            n_cpu_steps = int(min(200, max(4, round(12.0 / (probe_s / 2)))))
            rate, secs = cpu_reference_rate(512, n_cpu_steps, 1)
            unb = cpu_unbatched_rate(4, 100)
            cpu = {"value": rate, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                   "sample": f"batched CPU port, 512 of {ICS} ICs x {NX} cells x {n_cpu_steps} steps ({secs:.1f} s); "
                             f"the reference's own one-IC-at-a-time loop (oracle/ref_port.py, 4 ICs x 100 steps) "
                             f"reaches {unb:.3e} {UNIT}"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.precision == "fp32" else args.precision, "data": "synthetic",
            "config": {"workload": WORKLOAD, "l2": "flushed (256 MiB write) between timed steps",
                       "launches_per_step": launches / K, "wall_s_incl_flush": wall},
            "e2e": {"value": max(e2e_value, e2e_zero_copy), "unit": UNIT, "h2d_bytes_per_step": state_bytes,
                    "d2h_bytes_per_step": state_bytes,
                    "copy_engine": e2e_value, "zero_copy": e2e_zero_copy,
                    "note": "HybridSolver.step_pinned on pinned host state every step; `copy_engine` = cudaMemcpyAsync "
                            "H2D + kernel + D2H, `zero_copy` = the kernel reads/writes the pinned buffers over PCIe"},
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks.summary(),
            "tensor_path": tensor_path,
        }
        line["config"]["precision"] = args.precision
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)
    return 0


def run_side_workload(args):
    """Secondary, single-GPU measurements (not the driver's bench line): BASELINE.json
    configs[2] shape per GPU (c3: nx=1024, radius 2, window tiles + FFT field solve) and
    configs[4] (c5: classical solver alone at 2^24 cells, HBM-bound)."""
    if not torch.cuda.is_available():
        raise SystemExit("needs a CUDA device")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    from gnn_plasma_flux_b200 import BaselineSolver, FluxGNN, HybridSolver, MODEL_CONFIG, _lib
    from oracle import ref_port as P
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    K, W = args.steps, max(args.warmup, 3)
    if args.workload == "c5":
        nx, B = 1 << 24, args.batch or 1
        dt = 0.2 * (2 * np.pi / nx) ** 2 / 1e-3          # explicit viscosity: nu*dt/dx^2 = 0.2 (0.02*dx would be unstable)
        grid = P.Grid(nx=nx, dt=dt, nu=1e-3)
        ic = P.stable_initial_condition(grid, 0)
        state = torch.from_numpy(np.repeat(ic[None], B, 0)).to(dev)
        sol = BaselineSolver(nx=nx, dt=dt, nu=1e-3, device=dev)
        step = lambda st, n: sol.rollout(st, n)[0]
        name = f"C5 classical solver alone: {B} x 2^24 cells, nu=1e-3 (state {B * 3 * nx * 4 >> 20} MiB > L2, no flush needed)"
        flop = None
    else:
        nx, B, r = 1024, args.batch or 2048, 2
        dt = 3e-4
        grid = P.Grid(nx=nx, dt=dt)
        base = np.stack([P.stable_initial_condition(grid, s) for s in range(64)])
        state = torch.from_numpy(np.tile(base, (B // 64 + 1, 1, 1))[:B].copy()).to(dev)
        weights = P.init_weights(0)
        model = FluxGNN(**MODEL_CONFIG)
        model.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()})
        sol = HybridSolver(None, r, nx=nx, dt=dt, device=dev, graph_radius=r, model=model.to(dev))
        step = lambda st, n: sol.rollout(st, n)[0]
        name = f"C3 shape per GPU (scaled): {B} ICs x 1024 cells, radius 2 (state {B * 3 * nx * 4 >> 20} MiB)"
        flop = FLOP_PER_CELL_EXECUTED
    for _ in range(W):
        state = step(state, 1)
    torch.cuda.synchronize(dev)
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    state = step(state, K)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    assert torch.isfinite(state).all()
    rate = B * nx * K / (ms * 1e-3)
    line = {"metric": METRIC if flop else "baseline_rollout_cell_updates_per_sec", "value": rate, "unit": UNIT, "n_gpus": 1,
            "steps": K, "warmup": W, "ms_per_step": ms / K, "dtype": "f32", "data": "synthetic",
            "config": {"workload": name}, "gpu_launches": _lib.launch_count() - l0,
            "roofline": {"bound": "hbm" if not flop else "fp32-ffma", "achieved": rate * BYTES_PER_CELL / 1e9 if not flop else rate * flop / 1e12,
                         "peak": hbm_peak if not flop else None, "unit": "GB/s" if not flop else "TFLOP/s",
                         "frac": rate * BYTES_PER_CELL / 1e9 / hbm_peak if not flop else None, "traffic": None}}
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", choices=["c2", "c3", "c5"], default="c2",
                    help="c2 (default) is the driver's bench line; c3/c5 are secondary single-GPU measurements")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--precision", choices=["fp32", "tf32x3", "tf32"], default="fp32",
                    help="kernel of the headline number: fp32 = FP32-pipe FFMA kernel (default), tf32x3/tf32 = tcgen05 kernel")
    ap.add_argument("--no-tensor-path", action="store_true", help="skip the extra tensor-core measurements")
    args = ap.parse_args()
    if args.steps < 1:
        raise SystemExit("--steps must be >= 1")
    if args.workload != "c2":
        return run_side_workload(args)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
