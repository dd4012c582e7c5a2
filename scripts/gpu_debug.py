"""First-contact GPU check: print (not assert) parity errors for the main cases and a
rough timing of the C2 step.  Output goes to stdout; run under gpurun."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_plasma_flux_b200 import BaselineSolver, FluxGNN, HybridSolver, MODEL_CONFIG, _lib   # noqa: E402
from oracle import batched, ref_port as P                                                       # noqa: E402

weights = P.init_weights(0)
model = FluxGNN(**MODEL_CONFIG)
model.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()})
model = model.cuda().eval()
print("device", torch.cuda.get_device_name(0), flush=True)

for nx, dt, r, B in [(64, 5e-3, 1, 20), (64, 5e-3, 3, 20), (40, 5e-3, 2, 7), (1024, 3e-4, 2, 3), (200, 1e-3, 5, 3)]:
    grid = P.Grid(nx=nx, dt=dt)
    ics = np.stack([P.stable_initial_condition(grid, s) for s in range(B)])
    sol = HybridSolver(None, r, nx=nx, dt=dt, graph_radius=r, model=model)
    t0 = time.time()
    out = sol.step(ics)
    torch.cuda.synchronize()
    ref = batched.hybrid_step(weights, torch.from_numpy(ics), grid.x, grid.k, grid.dt, grid.dx, radius=r).numpy()
    print(f"step nx={nx} r={r} B={B}: rel err {P.rel_err(out, ref)}  finite={np.isfinite(out).all()}  {time.time()-t0:.3f}s", flush=True)

# timing: C2 shape
nx, B, r = 64, 4096, 3
grid = P.Grid(nx=nx, dt=1e-3)
ics = np.stack([P.stable_initial_condition(grid, s % 50) for s in range(B)])
sol = HybridSolver(None, r, nx=nx, dt=1e-3, graph_radius=r, model=model)
dev = torch.from_numpy(ics).cuda()
for steps in (1, 10, 100):
    sol.rollout(dev, steps)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out, _ = sol.rollout(dev, steps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    cu = B * nx * steps / (ms * 1e-3)
    print(f"C2 rollout steps={steps}: {ms:.3f} ms  {cu:.3e} cell-updates/s  ({cu*329216*1e-12:.1f} TFLOP/s split-W)", flush=True)
print("launches", _lib.launch_count())
