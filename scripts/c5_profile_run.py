"""BASELINE.json configs[4] for ncu: the classical solver alone at 2^24 cells, one warm-up rollout and one
profiled rollout of `steps` steps (the fused column kernel + row kernel per step).
    python scripts/c5_profile_run.py [steps] [batch]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_plasma_flux_b200 import BaselineSolver                                  # noqa: E402
from gnn_plasma_flux_b200.synthetic import stable_initial_conditions             # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
nx = 1 << 24
dt = 0.2 * (2 * np.pi / nx) ** 2 / 1e-3
sol = BaselineSolver(nx=nx, dt=dt, nu=1e-3, device="cuda")
state = stable_initial_conditions(sol, B)
state = sol.rollout(state, steps)[0]
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
state = sol.rollout(state, steps)[0]
e1.record()
torch.cuda.synchronize()
assert torch.isfinite(state).all()
print(f"{steps} steps, batch {B}: {e0.elapsed_time(e1) / steps:.4f} ms/step")
