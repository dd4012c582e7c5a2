"""Device time of the four-step field solve (fluxgnn_poisson_spectral) per column-pass variant:
FLUXGNN_FFT_TMA = 0 (plain kernels), 1 (TMA, dense tiles), 2 (TMA, 128-byte swizzled tiles where the tile rows are
128 bytes), and the cp.async-staged kernels.   python scripts/time_field_solve.py"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_plasma_flux_b200 import BaselineSolver                                  # noqa: E402

res = {}
for log2nx, B in ((22, 4), (23, 2), (24, 1)):
    nx = 1 << log2nx
    sol = BaselineSolver(nx=nx, device="cuda")
    n = 1.0 + 0.2 * torch.sin(torch.arange(nx, device="cuda") * (2 * np.pi * 5 / nx)).repeat(B, 1)
    for name, env in (("plain", {"FLUXGNN_FFT_TMA": "0"}), ("tma_dense", {"FLUXGNN_FFT_TMA": "1"}),
                      ("tma_default", {"FLUXGNN_FFT_TMA": "2"}), ("cp_async_staged", {"FLUXGNN_FFT_TMA": "0", "FLUXGNN_FFT_STAGING": "1"})):
        os.environ.pop("FLUXGNN_FFT_STAGING", None)
        os.environ.update(env)
        for _ in range(3):
            sol.solve_poisson(n)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            sol.solve_poisson(n)
        e1.record()
        torch.cuda.synchronize()
        res[f"2^{log2nx}x{B}_{name}"] = round(e0.elapsed_time(e1) / 20 * 1e3, 1)
print(json.dumps(res, indent=1))
