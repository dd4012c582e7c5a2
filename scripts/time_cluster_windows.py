"""Window tiles of the FP32-pipe kernel with and without thread-block clusters (C3 shape: nx = 1024, radius 2;
C4 slab shape: radius 3): device time per step for FLUXGNN_CLUSTER = 1..4 and for the library's own choice.
    python scripts/time_cluster_windows.py [ICs] [steps]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_plasma_flux_b200 import HybridSolver                                              # noqa: E402
from gnn_plasma_flux_b200.synthetic import seeded_model, stable_initial_conditions          # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda", 0)
model = seeded_model(0, dev)
res = {}
for name, nx, r, dt in (("c3_nx1024_r2", 1024, 2, 3e-4), ("nx4096_r3", 4096, 3, 7.5e-5)):
    sol = HybridSolver(None, r, nx=nx, dt=dt, device=dev, graph_radius=r, model=model)
    b = B if nx == 1024 else B // 4
    state = stable_initial_conditions(sol.baseline, b, distinct=64)
    out = torch.empty_like(state)
    ref = None
    for c in ("1", "2", "3", "4", "auto"):
        if c == "auto":
            os.environ.pop("FLUXGNN_CLUSTER", None)
        else:
            os.environ["FLUXGNN_CLUSTER"] = c
        sol.rollout(state, 1, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            sol.rollout(state, 1, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        if ref is None:
            ref = out.clone()
        res[f"{name}_cluster{c}"] = {"ms_per_step": ms, "cell_updates_per_s": b * nx / (ms * 1e-3),
                                     "bit_identical_to_cluster1": bool(torch.equal(out, ref))}
print(json.dumps(res, indent=1))
