"""Wall time of the reference-style calls for ONE initial condition of 64 cells (host numpy in and out)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_plasma_flux_b200 import HybridSolver, build_chain_graph                 # noqa: E402
from gnn_plasma_flux_b200.synthetic import seeded_model                          # noqa: E402

model = seeded_model(0, torch.device("cuda"))
sol = HybridSolver(None, 1, nx=64, dt=5e-3, device="cuda", model=model)
state = sol.baseline.initial_condition(seed=0)


def wall(fn, reps=200):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e6


print(f"solver.step(numpy [3,64])              {wall(lambda: sol.step(state)):8.1f} us per call")
print(f"solver.run(numpy, n_steps=50)          {wall(lambda: sol.run(state, n_steps=50), 50):8.1f} us per call")
dev_state = torch.from_numpy(state[None]).cuda()
print(f"solver.rollout(cuda [1,3,64], 1)       {wall(lambda: sol.rollout(dev_state, 1)):8.1f} us per call")
x = torch.as_tensor(sol.baseline.x, dtype=torch.float32)
nf, ei = build_chain_graph(state, x, device="cuda")
with torch.no_grad():
    print(f"model(node_features, edge_index)       {wall(lambda: model(nf, ei)):8.1f} us per call")
    print(f"build_chain_graph + model              {wall(lambda: model(*build_chain_graph(state, x, device='cuda'))):8.1f} us per call")
