"""Small drivers for ncu captures: a few launches of one workload after a warm-up of the same shape.
    python scripts/profile_run.py c2 [precision]     4096 ICs x 64 cells, radius 3 (BASELINE configs[1])
    python scripts/profile_run.py c3 [precision]     8192 ICs x 1024 cells, radius 2 (configs[2], the per-GPU share at 8 GPUs)
    python scripts/profile_run.py c5 [batch]         classical solver, 2^24 cells, 4-step rollout (configs[4])
    python scripts/profile_run.py dist               distributed field solve, 8 virtual ranks x 2^21 cells x batch 8 (configs[3])
    python scripts/profile_run.py generic            FluxGNN(4, 64, 3) hybrid step, 4096 ICs x 64 cells
    python scripts/profile_run.py latency            latency mode: 1 IC x 64 cells x 50 steps (the reference's timing protocol)
    python scripts/profile_run.py slabscan           distributed prefix-sum field solve, 8 virtual ranks x 2^21 cells x batch 8"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_plasma_flux_b200 import BaselineSolver, FluxGNN, HybridSolver                        # noqa: E402
from gnn_plasma_flux_b200.synthetic import seeded_model, stable_initial_conditions          # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "c2"
arg = sys.argv[2] if len(sys.argv) > 2 else None
dev = torch.device("cuda", 0)
if what in ("c2", "c3"):
    nx, r, dt, B = (64, 3, 1e-3, 4096) if what == "c2" else (1024, 2, 3e-4, 8192)
    sol = HybridSolver(None, r, nx=nx, dt=dt, device=dev, graph_radius=r, model=seeded_model(0, dev), precision=arg or "fp32")
    state = stable_initial_conditions(sol.baseline, B, distinct=64)
    for _ in range(5):
        state, _ = sol.rollout(state, 1)
elif what == "c5":
    nx = 1 << 24
    sol = BaselineSolver(nx=nx, dt=0.2 * (2 * np.pi / nx) ** 2 / 1e-3, nu=1e-3, device=dev)
    state = stable_initial_conditions(sol, int(arg or 1))
    for _ in range(2):
        state = sol.rollout(state, 4)[0]
elif what == "dist":
    from gnn_plasma_flux_b200.domain import DistributedFieldSolve, solve_emulated
    G, S, B = 8, 1 << 21, 8
    nx = G * S
    n = 1.0 + 0.1 * torch.sin(torch.arange(nx, device=dev, dtype=torch.float32) * (2 * np.pi * 3 / nx)).repeat(B, 1)
    E = torch.zeros_like(n)
    solvers = [DistributedFieldSolve(nx, 2 * np.pi, rk, G, dev) for rk in range(G)]
    for _ in range(2):
        solve_emulated(solvers, [n[:, k * S:(k + 1) * S] for k in range(G)], [E[:, k * S:(k + 1) * S] for k in range(G)])
    state = E
elif what == "generic":
    torch.manual_seed(0)
    model = FluxGNN(4, 64, 3).to(dev).eval()
    sol = HybridSolver(None, 1, nx=64, dt=1e-3, device=dev, model=model)
    state = stable_initial_conditions(sol.baseline, 4096, distinct=64)
    for _ in range(3):
        state, _ = sol.rollout(state, 1)
elif what == "latency":
    sol = HybridSolver(None, 1, nx=64, dt=5e-3, device=dev, model=seeded_model(0, dev))
    state = stable_initial_conditions(sol.baseline, 1)
    for _ in range(3):
        state, _ = sol.rollout(state, 50)
elif what == "slabscan":
    from gnn_plasma_flux_b200.domain import DistributedScanSolve, scan_solve_emulated
    G, S, B = 8, 1 << 21, 8
    nx = G * S
    n = 1.0 + 0.1 * torch.sin(torch.arange(nx, device=dev, dtype=torch.float32) * (2 * np.pi * 3 / nx)).repeat(B, 1)
    E = torch.zeros_like(n)
    solvers = [DistributedScanSolve(nx, 2 * np.pi, rk, G, dev) for rk in range(G)]
    for _ in range(2):
        scan_solve_emulated(solvers, [n[:, k * S:(k + 1) * S] for k in range(G)], [E[:, k * S:(k + 1) * S] for k in range(G)])
    state = E
else:
    raise SystemExit(__doc__)
torch.cuda.synchronize()
print("ok", what, float(state.abs().max()))
