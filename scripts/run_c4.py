"""BASELINE.json configs[3]: ONE large periodic grid, batch 8, radius 3, domain-decomposed over the
GPUs of one box (halo exchange + all-gathered density + replicated FFT field solve, NCCL).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/run_c4.py [--log2-cells-per-gpu 21] [--batch 8] [--steps 5] [--precision fp32|tf32x3]

Rank 0 first checks one decomposed step against the undivided single-GPU solver (bit-exact), then all
ranks time `steps` steps (barrier + synchronize on both sides, max over ranks) and rank 0 prints a JSON line.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_plasma_flux_b200 import BaselineSolver, HybridSolver                              # noqa: E402
from gnn_plasma_flux_b200.domain import DomainDecomposedHybridSolver, TorchDistComm        # noqa: E402
from gnn_plasma_flux_b200.synthetic import seeded_model, stable_initial_conditions          # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--log2-cells-per-gpu", type=int, default=21)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--radius", type=int, default=3)
ap.add_argument("--precision", default="fp32")
ap.add_argument("--no-check", action="store_true")
args = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
comm = TorchDistComm()

S = 1 << args.log2_cells_per_gpu
nx = S * world
dt = 0.02 * (2 * np.pi / nx)
model = seeded_model(0, dev)
full = stable_initial_conditions(BaselineSolver(nx=nx, dt=dt, device=dev), args.batch).cpu().numpy()   # same on every rank
local_state = torch.from_numpy(np.ascontiguousarray(full[..., rank * S:(rank + 1) * S])).to(dev)
sol = DomainDecomposedHybridSolver(model, nx, dt=dt, graph_radius=args.radius, rank=rank, world=world, device=dev,
                                   precision=args.precision)

check = None
if not args.no_check:
    one = sol.step(local_state, comm)
    gathered = comm.all_gather(one)                                                          # [world,B,3,S]
    if rank == 0:
        whole = HybridSolver(None, args.radius, nx=nx, dt=dt, device=dev, graph_radius=args.radius, model=model,
                             precision=args.precision)
        ref, _ = whole.rollout(torch.from_numpy(full).to(dev), 1)
        got = torch.cat(list(gathered.unbind(0)), dim=-1)
        check = {"bit_exact_vs_undivided": bool(torch.equal(got, ref)),
                 "max_abs_diff": float((got - ref).abs().max())}
        del whole, ref, got
    del gathered, one
    torch.cuda.empty_cache()

state = local_state
for _ in range(args.warmup):
    state = sol.step(state, comm)
dist.barrier()
torch.cuda.synchronize(dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    state = sol.step(state, comm)
e1.record()
dist.barrier()
torch.cuda.synchronize(dev)
ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
finite = torch.tensor([float(torch.isfinite(state).all())], device=dev)
dist.all_reduce(finite, op=dist.ReduceOp.MIN)
if rank == 0:
    t = float(ms.item())
    print(json.dumps({
        "metric": "hybrid_rollout_cell_updates_per_sec", "value": args.batch * nx * args.steps / (t * 1e-3),
        "unit": "cell-updates/s", "n_gpus": world, "steps": args.steps, "ms_per_step": t / args.steps,
        "scaling": "weak", "dtype": "f32" if args.precision == "fp32" else args.precision,
        "config": {"workload": f"C4 single grid: batch {args.batch} x {nx} cells (2^{args.log2_cells_per_gpu} per GPU), "
                               f"radius {args.radius}, domain-decomposed: halo {sol.halo} cells/side + all-gathered "
                               f"density + replicated FFT field solve"},
        "finite": bool(finite.item()), "check": check}), flush=True)
dist.destroy_process_group()
