"""Multi-GPU check of the peer-memory step (torchrun, one process per GPU):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/check_peer_step.py
Every rank advances its slab (1) with the collective step (NCCL halo exchange + all-gather, field_solve="scan"),
(2) with step_peer over symmetric memory, (3) with advance(graph=True), (4) with step_peer while the ranks' streams are
stalled at different steps (race hunt); all must agree bit for bit, and the first
step's n', u' must equal the undivided solver's.  Prints device time per step of each variant."""
import json, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_plasma_flux_b200 import BaselineSolver, HybridSolver
from gnn_plasma_flux_b200.domain import (DomainDecomposedBaselineSolver, DomainDecomposedHybridSolver, SymmetricMemoryFabric,
                                         TorchDistComm)
from gnn_plasma_flux_b200.synthetic import seeded_model, stable_initial_conditions

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
comm, fabric = TorchDistComm(), SymmetricMemoryFabric()
out = {}
for kind, nx, B, steps in (("baseline", 1 << 24, 1, 40), ("hybrid_fp16x3", 1 << 21, 8, 10)):
    S = nx // world
    if kind == "baseline":
        dt = 0.2 * (2 * np.pi / nx) ** 2 / 1e-3
        whole = BaselineSolver(nx=nx, dt=dt, nu=1e-3, device=dev, field_solve="spectral")
        make = lambda fab: DomainDecomposedBaselineSolver(nx, dt=dt, nu=1e-3, rank=rank, world=world, device=dev,
                                                          field_solve="scan", fabric=fab)
    else:
        model = seeded_model(0, dev)
        dt = 0.02 * (2 * np.pi / nx)
        whole = HybridSolver(None, 3, nx=nx, dt=dt, device=dev, graph_radius=3, model=model, precision="fp16x3")
        make = lambda fab: DomainDecomposedHybridSolver(model, nx, dt=dt, graph_radius=3, rank=rank, world=world, device=dev,
                                                        precision="fp16x3", field_solve="scan", fabric=fab)
    full = stable_initial_conditions(whole.baseline if kind != "baseline" else whole, B)
    local = full[..., rank * S:(rank + 1) * S].contiguous()
    want1 = whole.rollout(full, 1)[0][..., rank * S:(rank + 1) * S]
    res, ms, certified = {}, {}, {}
    for variant in ("collective", "peer", "peer_graph", "peer_skewed"):
        sol = make(None if variant == "collective" else fabric)
        counter = [0]

        def run(state, n):
            if variant == "collective":
                for _ in range(n):
                    state = sol.step(state, comm)
                return state
            if variant == "peer_skewed":                   # race hunt: the ranks' streams stall at different steps
                for _ in range(n):
                    counter[0] += 1
                    if (counter[0] * 7 + rank * 3) % 5 == 0:
                        torch.cuda._sleep(3_000_000)
                    state = sol.step_peer(state)
                return state
            return sol.advance(state, n, graph=variant == "peer_graph")

        first = run(local, 1).clone()
        assert torch.equal(first[:, :2], want1[:, :2]), (kind, variant, "first step n', u' vs undivided")
        state = run(first, steps - 1)                       # warm-up + graph capture
        res[variant] = state.clone()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        state = run(state, steps)
        e1.record(); torch.cuda.synchronize(); dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms[variant] = float(t)
        assert torch.isfinite(state).all()
        bad = sol.first_uncertified(state, comm)          # (the hybrid step's kinked n' certifies on long grids only)
        assert bad is None or kind != "baseline", (kind, variant, bad)
        certified[variant] = bad is None
    for variant in ("peer", "peer_graph", "peer_skewed"):
        assert torch.equal(res[variant], res["collective"]), (kind, variant, "differs from the collective step")
    out[kind] = {"nx": nx, "batch": B, "ranks": world, "ms_per_step": ms,
                 "cell_updates_per_sec": {k: B * nx / (v * 1e-3) for k, v in ms.items()}, "bit_identical": True,
                 "all_fields_certified": certified}
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()
