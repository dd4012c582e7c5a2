"""First-contact check of the tensor-core path: parity vs the batched oracle + timing."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_plasma_flux_b200 import FluxGNN, HybridSolver, MODEL_CONFIG, _lib   # noqa: E402
from oracle import batched, ref_port as P                                    # noqa: E402

weights = P.init_weights(0)
model = FluxGNN(**MODEL_CONFIG)
model.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()})
model = model.cuda().eval()

for prec in ("tf32x3", "tf32"):
    for nx, dt, r, B in [(64, 5e-3, 1, 4), (64, 5e-3, 3, 20), (128, 1e-3, 2, 3), (32, 5e-3, 2, 9), (1024, 3e-4, 2, 3)]:
        grid = P.Grid(nx=nx, dt=dt)
        ics = np.stack([P.stable_initial_condition(grid, s) for s in range(B)])
        sol = HybridSolver(None, r, nx=nx, dt=dt, graph_radius=r, model=model, precision=prec)
        ref32 = HybridSolver(None, r, nx=nx, dt=dt, graph_radius=r, model=model)
        out = sol.step(ics)
        torch.cuda.synchronize()
        ref = batched.hybrid_step(weights, torch.from_numpy(ics), grid.x, grid.k, grid.dt, grid.dx, radius=r).numpy()
        o32 = ref32.step(ics)
        # flux-level error: (n' - n) carries c * dF
        dn_ref = (ref[:, 0] - ics[:, 0]).astype(np.float64)
        dn = (out[:, 0] - ics[:, 0]).astype(np.float64)
        ferr = np.abs(dn - dn_ref).max() / np.abs(dn_ref).max()
        print(f"{prec} step nx={nx} r={r} B={B}: state rel err {P.rel_err(out, ref)}  flux-level rel err {ferr:.2e}  "
              f"fp32-kernel state err {P.rel_err(o32, ref)} finite={np.isfinite(out).all()}", flush=True)

nx, B, r = 64, 4096, 3
grid = P.Grid(nx=nx, dt=1e-3)
ics = np.stack([P.stable_initial_condition(grid, s % 50) for s in range(B)])
dev = torch.from_numpy(ics).cuda()
for prec in ("fp32", "tf32x3", "tf32"):
    sol = HybridSolver(None, r, nx=nx, dt=1e-3, graph_radius=r, model=model, precision=prec)
    for steps in (1, 20):
        sol.rollout(dev, steps)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out, _ = sol.rollout(dev, steps)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"{prec} C2 rollout steps={steps}: {ms:.3f} ms  {B * nx * steps / (ms * 1e-3):.3e} cell-updates/s", flush=True)
