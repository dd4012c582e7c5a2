"""C3 shape (1024 cells, radius 2) with and without packed remainder windows (FLUXGNN_NO_PACK=1), fp32 and fp16x3."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_plasma_flux_b200 import BaselineSolver, HybridSolver
from gnn_plasma_flux_b200.synthetic import seeded_model, stable_initial_conditions

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
m = seeded_model(0, "cuda")
for precision in ("fp32", "fp16x3"):
    sol = HybridSolver(None, 2, nx=1024, dt=3e-4, model=m, graph_radius=2, precision=precision)
    st = stable_initial_conditions(sol.baseline, B)
    for nopack in ("1", "0", "1", "0"):                     # alternate: the tensor modes run into the power cap
        os.environ["FLUXGNN_NO_PACK"] = nopack
        steps = 2 if precision == "fp32" else 20
        sol.rollout(st, 1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out, _ = sol.rollout(st, steps)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        print(f"{precision} B={B} nx=1024 r=2 packing={'off' if nopack == '1' else 'on'}: {ms:.3f} ms/step -> {B*1024/(ms*1e-3):.4e} cell-updates/s")
