"""World-size-2 gloo test (CPU) of the domain-decomposition host logic: ring halo exchange,
slab indexing, global x feature, density all-gather and slab extraction.  The local slab step
and the field solve are replaced by the CPU oracle (a checker standing in for the CUDA calls),
so the distributed result must equal the oracle on the undivided grid."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import batched, ref_port as P


class _Model:
    num_layers = 4


def _oracle_slab_fn(weights, grid, radius, x_ext):
    """Hybrid update of an extended slab with the batched oracle (periodic roll on the extended
    array is wrong only within `halo` cells of its ends, which are ghosts)."""
    def fn(ext):
        H = 4 * radius + 1
        fl = batched.edge_fluxes(weights, ext, x_ext, radius, hops=1)
        m = ext.shape[-1]
        face = 0.5 * (fl[:, :m] + fl[:, m:])
        c, dt32 = float(np.float32(grid.dt / grid.dx)), float(np.float32(grid.dt))
        n, u, E = ext[:, 0], ext[:, 1], ext[:, 2]
        n_new = n - c * (face - torch.roll(face, 1, dims=-1))
        fu = 0.5 * u * u
        u_new = (u - c * (fu - torch.roll(fu, 1, dims=-1))) + dt32 * E
        out = torch.stack([n_new, u_new, torch.zeros_like(n_new)], dim=1)
        return out[..., H:m - H].contiguous()
    return fn


def _worker(rank, world, port, nx, radius, steps, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gnn_plasma_flux_b200.domain import DomainDecomposedHybridSolver, TorchDistComm, split_slabs
        torch.set_num_threads(2)
        weights = P.init_weights(0)
        grid = P.Grid(nx=nx, dt=1e-3)
        ics = torch.from_numpy(np.stack([P.stable_initial_condition(grid, s) for s in range(2)]))
        k = torch.as_tensor(grid.k)
        sol = DomainDecomposedHybridSolver(_Model(), nx, dt=1e-3, graph_radius=radius, rank=rank, world=world,
                                           device="cpu", slab_fn=None, field_fn=lambda n: batched.poisson(n, k))
        sol._slab_fn = _oracle_slab_fn(weights, grid, radius, sol.x_ext)
        np.testing.assert_array_equal(
            sol.x_ext.numpy(), grid.x[(rank * sol.owned - sol.halo + np.arange(sol.owned + 2 * sol.halo)) % nx].astype(np.float32))
        comm = TorchDistComm()
        local = split_slabs(ics, world)[rank]
        ref = ics
        for _ in range(steps):
            local = sol.step(local, comm)
            ref = batched.hybrid_step(weights, ref, grid.x, grid.k, grid.dt, grid.dx, radius=radius)
        want = split_slabs(ref, world)[rank]
        err = P.rel_err(local.numpy(), want.numpy()).max()
        out[rank] = float(err)
    finally:
        dist.destroy_process_group()


def test_two_rank_domain_decomposition_matches_global_oracle():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(2, port, 256, 2, 3, out), nprocs=2, join=True)
    assert out[0] < 1e-6 and out[1] < 1e-6, dict(out)
