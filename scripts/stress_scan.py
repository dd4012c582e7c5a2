"""Race hunt for the scan-solve rollout: many repetitions at large batch, first-step n', u' must equal the FFT path's
bit for bit (the field is not involved in the first step) and every multi-step rollout must be certified."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_plasma_flux_b200 import BaselineSolver                                  # noqa: E402
from gnn_plasma_flux_b200.synthetic import stable_initial_conditions             # noqa: E402

for nx, B, trials in [(1 << 20, 16, 30), (1 << 22, 8, 20), (1 << 24, 8, 4), (1 << 18, 40, 30), (1 << 24, 1, 10)]:
    dt = 0.2 * (2 * np.pi / nx) ** 2 / 1e-3
    sol = BaselineSolver(nx=nx, dt=dt, nu=1e-3, device="cuda")
    state = stable_initial_conditions(sol, B)
    ref1 = sol.rollout(state, 1, field_solve="spectral")[0]
    nbad, uncert = 0, 0
    for _ in range(trials):
        out = sol.rollout(state, 1, field_solve="scan")[0]
        nbad += int((out[:, :2] != ref1[:, :2]).sum())
        sol.rollout(state, 7, field_solve="auto")
        uncert += sol.last_field_solve != "scan"
    print(f"nx={nx} B={B}: {trials} trials, cells differing in n', u' after one step: {nbad}, uncertified rollouts: {uncert}")
    del state, ref1, out
    torch.cuda.empty_cache()
