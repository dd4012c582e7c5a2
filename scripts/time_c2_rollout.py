"""BASELINE.json configs[1] as ONE call: 4096 ICs x 64 cells, radius 3, 1000 steps in a single persistent launch
(state never leaves the SM between steps), timed with CUDA events.  usage: python scripts/time_c2_rollout.py [precision ...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_plasma_flux_b200 import HybridSolver                                               # noqa: E402
from gnn_plasma_flux_b200.synthetic import seeded_model, stable_initial_conditions          # noqa: E402

B, NX, R, STEPS = 4096, 64, 3, 1000
dev = torch.device("cuda", 0)
model = seeded_model(0, dev)
for prec in (sys.argv[1:] or ["fp32", "fp16x3"]):
    sol = HybridSolver(None, R, nx=NX, dt=1e-3, device=dev, graph_radius=R, model=model, precision=prec)
    state = stable_initial_conditions(sol.baseline, B, distinct=64)
    sol.rollout(state, 10)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    final, _ = sol.rollout(state, STEPS)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(json.dumps({"workload": f"C2 rollout, {B} ICs x {NX} cells x {STEPS} steps in one launch", "precision": prec,
                      "ms_total": ms, "ms_per_step": ms / STEPS, "cell_updates_per_s": B * NX * STEPS / (ms * 1e-3),
                      "finite": bool(torch.isfinite(final).all())}))
