"""Latency mode (cluster of 8 CTAs per tile, csrc/hybrid_latency_kernel.cu) against the tile kernel: bit-identity and
wall time of the reference's timing protocol (1 IC x 64 cells x 50 steps) and of BASELINE.json configs[0]."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_plasma_flux_b200 import HybridSolver                                    # noqa: E402
from gnn_plasma_flux_b200.synthetic import seeded_model, stable_initial_conditions   # noqa: E402

model = seeded_model(0, torch.device("cuda"))


def run(sol, state, steps, mode, record):
    os.environ["FLUXGNN_LATENCY"] = mode
    out = sol.rollout(state, steps, record_every=record)
    torch.cuda.synchronize()
    return out


bad = 0
for nx, B, r, steps in [(64, 1, 1, 1), (64, 1, 3, 7), (64, 20, 1, 30), (64, 5, 2, 4), (32, 7, 3, 5), (128, 3, 2, 6), (40, 4, 1, 5),
                        (36, 3, 2, 4), (100, 2, 4, 3), (64, 70, 3, 3)]:
    sol = HybridSolver(None, r, nx=nx, dt=1e-3, device="cuda", graph_radius=r, model=model)
    state = stable_initial_conditions(sol.baseline, B)
    a, ta = run(sol, state, steps, "0", 1)
    b, tb = run(sol, state, steps, "1", 1)
    same = bool(torch.equal(a, b)) and bool(torch.equal(ta, tb))
    bad += not same
    print(f"nx={nx} B={B} r={r} steps={steps}: bit-identical {same}, max dev {float((ta - tb).abs().max()):.2e}", flush=True)
print("mismatches:", bad)

for name, nx, B, r, steps in [("reference protocol: 1 IC x 64 cells x 50 steps", 64, 1, 1, 50), ("C1: 20 ICs x 64 x 30", 64, 20, 1, 30),
                              ("4 ICs x 128 x 30, radius 3", 128, 4, 3, 30)]:
    sol = HybridSolver(None, r, nx=nx, dt=5e-3 if nx == 64 else 1e-3, device="cuda", graph_radius=r, model=model)
    state = stable_initial_conditions(sol.baseline, B)
    for mode in ("0", "1"):
        os.environ["FLUXGNN_LATENCY"] = mode
        for _ in range(3):
            sol.rollout(state, steps)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            sol.rollout(state, steps)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / 10 * 1e3
        print(f"{name}: latency mode {mode}: {ms:.3f} ms per rollout, {ms / steps * 1e3:.1f} us per step")
os.environ.pop("FLUXGNN_LATENCY")
