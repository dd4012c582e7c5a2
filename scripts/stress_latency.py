"""Race hunt for the latency mode (clusters of 8 CTAs exchanging activations through distributed shared memory): many
rollouts of random small shapes, each compared bit for bit with the tile kernel."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_plasma_flux_b200 import HybridSolver                                        # noqa: E402
from gnn_plasma_flux_b200.synthetic import seeded_model, stable_initial_conditions   # noqa: E402

model = seeded_model(0, torch.device("cuda"))
rng = np.random.RandomState(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
solvers, bad, runs = {}, 0, 0
for trial in range(int(sys.argv[2]) if len(sys.argv) > 2 else 150):
    nx = int(rng.choice([16, 32, 36, 40, 64, 64, 64, 96, 128]))
    r = int(rng.randint(1, 5))
    B = int(rng.choice([1, 1, 2, 3, 5, 8, 20, 33, 64, 150]))
    steps = int(rng.choice([1, 2, 5, 17, 50]))
    sol = solvers.setdefault((nx, r), HybridSolver(None, r, nx=nx, dt=1e-3, device="cuda", graph_radius=r, model=model))
    state = stable_initial_conditions(sol.baseline, B, first_seed=trial)
    os.environ["FLUXGNN_LATENCY"] = "0"
    want = sol.rollout(state, steps, record_every=max(1, steps // 3))
    os.environ["FLUXGNN_LATENCY"] = "1"
    for rep in range(3):
        got = sol.rollout(state, steps, record_every=max(1, steps // 3))
        runs += 1
        if not (torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])):
            bad += 1
            print(f"MISMATCH trial {trial} rep {rep}: nx={nx} r={r} B={B} steps={steps} max dev {float((got[0] - want[0]).abs().max()):.3e}",
                  flush=True)
print(f"{runs} latency-mode rollouts compared with the tile kernel: {bad} mismatches")
