import os, sys
sys.path.insert(0, os.getcwd())
import torch
from gnn_plasma_flux_b200 import HybridSolver
from gnn_plasma_flux_b200.synthetic import seeded_model, stable_initial_conditions
dev = torch.device("cuda", 0)
sol = HybridSolver(None, 3, nx=64, dt=1e-3, device=dev, graph_radius=3, model=seeded_model(0, dev))
st = stable_initial_conditions(sol.baseline, 4096, distinct=64)
sol.rollout(st, 1); torch.cuda.synchronize()
print("---- timed launch (1 step, 4096 ICs)")
sol.rollout(st, 1); torch.cuda.synchronize()
