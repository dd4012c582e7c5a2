"""Launch loop vs CUDA-graph replay of the window-mode rollout (nx = 1024, radius 2) at small batch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gnn_plasma_flux_b200 import HybridSolver
from gnn_plasma_flux_b200.synthetic import seeded_model, stable_initial_conditions

dev = torch.device("cuda", 0)
for prec in ("fp32", "fp16x3"):
    for B in (1, 8, 64):
        sol = HybridSolver(None, 2, nx=1024, dt=3e-4, device=dev, graph_radius=2, model=seeded_model(0, dev), precision=prec)
        st = stable_initial_conditions(sol.baseline, B)
        res = {}
        for name, fn in (("launch loop", lambda: sol.rollout(st, 200)[0]), ("graph replay", lambda: sol.rollout_graphed(st, 200, chunk=20))):
            fn(); torch.cuda.synchronize()
            t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize()
            res[name] = (time.perf_counter() - t0) / 200 * 1e6
        print(f"{prec:7s} B={B:3d}: launch loop {res['launch loop']:8.1f} us/step, graph replay {res['graph replay']:8.1f} us/step")
