"""The reference's timing protocol (scripts/evaluation/benchmark_timing.py:25-26,34-35,63-72: 50-step rollout
from `initial_condition(seed=42)` at nx = 64, 10 runs, wall clock including the host round trip) run
through the drop-in classes: classical solver, hybrid solver (fp32 and fp16x3 kernels), PureGNN, PINN.
Random-init weights of the reference's architectures (no checkpoints travel with the repo)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_plasma_flux_b200 import BaselineSolver, HybridSolver, PINN, PureGNN          # noqa: E402
from gnn_plasma_flux_b200.synthetic import seeded_model                                # noqa: E402

N_RUNS, N_STEPS = 10, 50
dev = torch.device("cuda", 0)


def timed(fn):
    fn()                                                    # warm-up (lazy packing, attributes)
    ts = []
    for _ in range(N_RUNS):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return {"mean_ms": 1e3 * float(np.mean(ts)), "std_ms": 1e3 * float(np.std(ts))}


base = BaselineSolver(device=dev)
state0 = base.initial_condition(seed=42)
res = {"protocol": f"{N_STEPS}-step rollout of one IC (nx=64, seed 42), {N_RUNS} runs, wall clock, numpy in / numpy out"}
res["classical"] = timed(lambda: base.run(state0, n_steps=N_STEPS))
model = seeded_model(0, dev)
for prec in ("fp32", "fp16x3"):
    sol = HybridSolver(None, 3, device=dev, model=model, precision=prec)
    res[f"hybrid_{prec}"] = timed(lambda: sol.run(state0, n_steps=N_STEPS))
torch.manual_seed(0)
pg = PureGNN(input_dim=4, hidden_dim=128, num_layers=4).to(dev)
x = torch.from_numpy(base.x.astype(np.float32)).to(dev)
res["pure_gnn"] = timed(lambda: pg.rollout(torch.from_numpy(state0)[None].to(dev), x, N_STEPS).cpu().numpy())
pinn = PINN(input_dim=3 * 64, hidden_dim=256, num_layers=4).to(dev)


def pinn_rollout():
    s = torch.from_numpy(state0)[None].to(dev)
    for _ in range(N_STEPS):
        s = pinn(s)
    return s.cpu().numpy()


res["pinn"] = timed(pinn_rollout)
print(json.dumps(res, indent=1))
