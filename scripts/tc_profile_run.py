"""Small driver for ncu: a few launches of the tensor-core hybrid step at the C2 shape."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_plasma_flux_b200 import FluxGNN, HybridSolver, MODEL_CONFIG   # noqa: E402
from oracle import ref_port as P                                        # noqa: E402  (inputs only)

prec = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
weights = P.init_weights(0)
model = FluxGNN(**MODEL_CONFIG)
model.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()})
nx, B, r = 64, 4096, 3
grid = P.Grid(nx=nx, dt=1e-3)
ics = np.stack([P.stable_initial_condition(grid, s % 50) for s in range(B)])
sol = HybridSolver(None, r, nx=nx, dt=1e-3, graph_radius=r, model=model.cuda(), precision=prec)
dev = torch.from_numpy(ics).cuda()
for _ in range(5):
    dev, _ = sol.rollout(dev, 1)
torch.cuda.synchronize()
print("ok", float(dev.abs().max()))
