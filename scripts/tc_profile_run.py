"""Small driver for ncu: a few launches of a tensor-core hybrid step at the C2 shape (4096 ICs x 64 cells, radius 3).
usage: python scripts/tc_profile_run.py [fp16x3|fp16|bf16|tf32x3|tf32]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_plasma_flux_b200 import HybridSolver                                               # noqa: E402
from gnn_plasma_flux_b200.synthetic import seeded_model, stable_initial_conditions          # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "fp16x3"
dev = torch.device("cuda", 0)
sol = HybridSolver(None, 3, nx=64, dt=1e-3, device=dev, graph_radius=3, model=seeded_model(0, dev), precision=prec)
state = stable_initial_conditions(sol.baseline, 4096, distinct=64)
for _ in range(5):
    state, _ = sol.rollout(state, 1)
torch.cuda.synchronize()
print("ok", float(state.abs().max()))
