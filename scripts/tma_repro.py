import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from gnn_plasma_flux_b200 import BaselineSolver
log2nx = int(sys.argv[1]); B = int(sys.argv[2])
nx = 1 << log2nx
sol = BaselineSolver(nx=nx, device="cuda")
n = 1.0 + 0.2 * torch.sin(torch.arange(nx, device="cuda") * (2 * np.pi * 5 / nx)).repeat(B, 1)
os.environ["FLUXGNN_FFT_TMA"] = "0"
plain = sol.solve_poisson(n); torch.cuda.synchronize()
for mode in ("1", "2"):
    os.environ["FLUXGNN_FFT_TMA"] = mode
    got = sol.solve_poisson(n); torch.cuda.synchronize()
    print("mode", mode, "equal", bool(torch.equal(got, plain)), float((got - plain).abs().max()))
