"""Forward-with-save + backward timing of the trainable FluxGNN at the C2 shape (4096 ICs x 64 cells)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_plasma_flux_b200 import BaselineSolver
from gnn_plasma_flux_b200.autograd import ring_fluxes_with_grad
from gnn_plasma_flux_b200.synthetic import seeded_model, stable_initial_conditions
m = seeded_model(0, "cuda").train()
sol = BaselineSolver(nx=64, dt=1e-3)
for B in (64, 4096):
    st = stable_initial_conditions(sol, B).requires_grad_(True)
    x = torch.as_tensor(sol.x, dtype=torch.float32, device="cuda")
    for r in (1, 3):
        def step():
            for p in m.parameters(): p.grad = None
            fl = ring_fluxes_with_grad(m, st, x, r, 1)
            fl.square().mean().backward()
        for _ in range(3): step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): step()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"B={B} nx=64 radius={r}: forward+backward {ms:.3f} ms  -> {B*64/(ms*1e-3):.3e} cell-gradients/s "
              f"({B*64*3*327680/(ms*1e-3)/1e12:.1f} TFLOP/s of GEMM work)")
