"""Two training forward + backward passes at the C2 shape (4096 ICs x 64 cells, radius 3), nothing else: the command behind
the ncu launch list profiles/r2_train_launches.csv."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from gnn_plasma_flux_b200 import BaselineSolver
from gnn_plasma_flux_b200.autograd import ring_fluxes_with_grad
from gnn_plasma_flux_b200.synthetic import seeded_model, stable_initial_conditions
m = seeded_model(0, "cuda").train()
sol = BaselineSolver(nx=64, dt=1e-3)
st = stable_initial_conditions(sol, 4096).requires_grad_(True)
x = torch.as_tensor(sol.x, dtype=torch.float32, device="cuda")
for _ in range(2):
    for p in m.parameters(): p.grad = None
    fl = ring_fluxes_with_grad(m, st, x, 3, 1)
    fl.square().mean().backward()
torch.cuda.synchronize()
