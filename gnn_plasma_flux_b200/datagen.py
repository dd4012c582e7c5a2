"""Training-data generation on the device, drop-in for
scripts/training/generate_data.py:12-54 of the reference: the classical solver is rolled out
from `num_initial_conditions` seeded ICs and (state_t, F_n, state_next) triples are stored in
the same .npz layout.  The reference loops ICs one at a time on the host; here all ICs advance
in ONE batched device rollout (BaselineSolver.rollout with record_flux)."""
from __future__ import annotations

import os

import numpy as np
import torch

from .baseline_solver import BaselineSolver


def generate_dataset(nx=64, num_initial_conditions=20, steps_per_ic=30, dt=5e-3, t_end=1.0, nu=1e-3,
                     out_path="data/dataset.npz", *, device="cuda"):
    solver = BaselineSolver(nx=nx, dt=dt, t_end=t_end, nu=nu, device=device)
    ics = np.stack([solver.initial_condition(seed=ic) for ic in range(num_initial_conditions)])      # seeds as the reference
    state0 = torch.from_numpy(ics).to(solver.device)
    _, traj, flux = solver.rollout(state0, steps_per_ic, record_every=1, record_flux=True)
    states = torch.cat([state0.unsqueeze(0), traj], dim=0)                     # [T+1,B,3,nx]
    # reference order: IC-major, then time (np.concatenate over ICs of [steps,...] blocks)
    state_t = states[:-1].permute(1, 0, 2, 3).reshape(-1, 3, nx).cpu().numpy()
    state_next = states[1:].permute(1, 0, 2, 3).reshape(-1, 3, nx).cpu().numpy()
    flux_t = flux.permute(1, 0, 2).reshape(-1, nx).cpu().numpy()
    x = solver.x.astype(np.float32)
    if out_path:
        folder = os.path.dirname(out_path)
        if folder:
            os.makedirs(folder, exist_ok=True)
        np.savez(out_path, state_t=state_t, flux_t=flux_t, state_next=state_next, x=x, dt=dt, dx=solver.dx, nu=nu)
    return state_t, flux_t, state_next, x, dt, solver.dx, nu
