"""Synthetic inputs for benchmarks and long rollouts (product-side; no oracle involved).

* `seeded_model(seed)`: `torch.manual_seed(seed); FluxGNN(**MODEL_CONFIG)` -- the random-init weights
  of the named architecture (the construction order equals the reference's, so the tensors equal
  `torch.manual_seed(seed); src.flux_gnn.FluxGNN(4, 128, 4)`).
* `stable_initial_conditions`: the reference's random-mode IC family (src/baseline_solver.py:29-57)
  made safe for 1000-step rollouts (the reference's own ICs go non-finite after
  ~200 steps): 4 density sine modes and 2 velocity cosine modes, amplitude in [0, 0.1), mode number
  1..5, uniform phase, no white noise; E from the field solve of this package.
"""
from __future__ import annotations

import numpy as np
import torch

from .baseline_solver import BaselineSolver
from .config import MODEL_CONFIG
from .flux_gnn import FluxGNN


def seeded_model(seed: int = 0, device="cuda") -> FluxGNN:
    torch.manual_seed(seed)
    return FluxGNN(**MODEL_CONFIG).to(device).eval()


def stable_density_velocity(x: np.ndarray, seed: int):
    rng = np.random.RandomState(seed)
    n = np.full(x.shape[0], 1.0, dtype=np.float64)
    for _ in range(4):
        mode, amp, phase = rng.randint(1, 6), 0.1 * rng.rand(), 2 * np.pi * rng.rand()
        n += amp * np.sin(mode * x + phase)
    u = np.zeros(x.shape[0], dtype=np.float64)
    for _ in range(2):
        mode, amp, phase = rng.randint(1, 6), 0.1 * rng.rand(), 2 * np.pi * rng.rand()
        u += amp * np.cos(mode * x + phase)
    return n.astype(np.float32), u.astype(np.float32)


def stable_initial_conditions(solver: BaselineSolver, n_ics: int, first_seed: int = 0, distinct: int = 256) -> torch.Tensor:
    """[n_ics,3,nx] float32 on the solver's device.  `distinct` seeded ICs are generated and tiled; a
    small per-IC velocity offset makes every member of the ensemble different."""
    base = min(n_ics, distinct)
    nu = [stable_density_velocity(solver.x, first_seed + s) for s in range(base)]
    n = torch.from_numpy(np.stack([a for a, _ in nu])).to(solver.device)
    u = torch.from_numpy(np.stack([b for _, b in nu])).to(solver.device)
    E = solver.solve_poisson(n)
    state = torch.stack([n, u, E], dim=1)
    reps = (n_ics + base - 1) // base
    state = state.repeat(reps, 1, 1)[:n_ics].contiguous()
    off = np.random.RandomState(first_seed).uniform(-1e-2, 1e-2, size=(n_ics, 1)).astype(np.float32)
    state[:, 1] += torch.from_numpy(off).to(solver.device)
    return state
