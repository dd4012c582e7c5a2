"""World-size-2 gloo tests (CPU) of the domain-decomposition host logic: extended ping-pong state, ring halo
exchange, slab indexing, global x feature, and BOTH field-solve variants -- the all-gathered density with a
replicated solve, and the distributed solve (pairs of ICs as complex signals, four all-to-alls, rank-index
decimation).  The CUDA calls (slab step, the four stages of the distributed solve) are replaced by CPU
stand-ins built on the oracle (checkers standing in for the kernels), so the distributed result must equal
the oracle on the undivided grid."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import batched, ref_port as P


class _Model:
    num_layers = 4


def _oracle_slab_fn(weights, grid, radius, sol):
    """Hybrid update of an extended slab with the batched oracle (periodic roll on the extended
    array is wrong only within `halo` cells of its ends, which are ghosts); writes n', u' into the interior
    of the next extended buffer like fluxgnn_hybrid_slab_step_ld."""
    def fn(ext, nxt):
        H = sol.halo
        fl = batched.edge_fluxes(weights, ext, sol.x_ext, radius, hops=1)
        m = ext.shape[-1]
        face = 0.5 * (fl[:, :m] + fl[:, m:])
        c, dt32 = float(np.float32(grid.dt / grid.dx)), float(np.float32(grid.dt))
        n, u, E = ext[:, 0], ext[:, 1], ext[:, 2]
        n_new = n - c * (face - torch.roll(face, 1, dims=-1))
        fu = 0.5 * u * u
        u_new = (u - c * (fu - torch.roll(fu, 1, dims=-1))) + dt32 * E
        nxt[:, 0, H:m - H] = n_new[:, H:m - H]
        nxt[:, 1, H:m - H] = u_new[:, H:m - H]
    return fn


def _baseline_slab_fn(grid, sol):
    """Classical update of an extended slab with the batched oracle (one ghost cell is read)."""
    def fn(ext, nxt):
        H, m = sol.halo, ext.shape[-1]
        out = batched.baseline_step(ext, np.ones(m), grid.dt, grid.dx, grid.nu)     # k unused for n', u'
        nxt[:, :2, H:m - H] = out[:, :2, H:m - H]
    return fn


def cpu_field_stages():
    """CPU restatement of fluxgnn_poisson_dist_{pack,rank_dft,local,unpack} (include/fluxgnn.h) in complex128."""
    def pack(n_rows, z):
        B, S = n_rows.shape
        zz = z.view(-1, S, 2)
        zz.zero_()
        for b in range(B):
            zz[b // 2, :, b % 2] = n_rows[b] - 1.0

    def rank_dft(sol, src, dst, inverse):
        G, S, nx = sol.world, sol.S, sol.nx
        v = torch.view_as_complex(src.view(G, -1, 2).contiguous()).to(torch.complex128)     # [q][e]
        chunk = v.shape[1]
        jl = (sol.rank * chunk + torch.arange(chunk)) % S
        r = torch.arange(G)
        sign = 1.0 if inverse else -1.0
        tw = torch.exp(sign * 2j * np.pi * (jl[None, :] * r[:, None]).double() / nx)         # [r][e] W_nx^(sign jl r)
        wg = torch.exp(sign * 2j * np.pi * (r[:, None] * r[None, :]).double() / G)           # [r][q]
        if inverse:
            out = (wg @ (v * tw)) / G
        else:
            out = (wg @ v) * tw
        dst.view(G, -1, 2).copy_(torch.view_as_real(out.to(torch.complex64)))

    def local(sol, y):
        G, S, nx = sol.world, sol.S, sol.nx
        v = torch.view_as_complex(y.view(-1, S, 2)).to(torch.complex128)
        spec = torch.fft.fft(v, dim=-1)
        k = sol.rank + G * torch.arange(S)
        kk = torch.where(2 * k < nx, k, k - nx).double()
        mult = torch.zeros(S, dtype=torch.complex128)
        ok = (k != 0) & (2 * k != nx)
        mult[ok] = 1j * (sol.length / (2 * np.pi)) / kk[ok]
        y.view(-1, S, 2).copy_(torch.view_as_real(torch.fft.ifft(spec * mult, dim=-1).to(torch.complex64)))

    def unpack(e, E_rows):
        B, S = E_rows.shape
        ee = e.view(-1, S, 2)
        for b in range(B):
            E_rows[b] = ee[b // 2, :, b % 2]

    return {"pack": pack, "rank_dft": rank_dft, "local": local, "unpack": unpack}


def cpu_scan_stages():
    """CPU restatement of fluxgnn_scan_slab_{sums,field,certify} (include/fluxgnn.h) in float64.  Message per IC:
    [S, M1, D4, maxE, n_first0, n_first1, n_last0, n_last1]."""
    def sums(sol, n_rows, st):
        S = sol.S
        rho = (n_rows - 1.0).double()
        j = (sol.rank * S + torch.arange(S)).double()
        st["msg"][:, 0] = rho.sum(-1)
        st["msg"][:, 1] = (rho * j).sum(-1)
        st["msg"][:, 2:4] = st["cert"]                       # certificate sums of the field reconstructed last
        st["msg"][:, 4:6] = n_rows[:, :2].double()
        st["msg"][:, 6:8] = n_rows[:, -2:].double()

    def field(sol, n_rows, E_rows, msg_all, st):
        G, S = sol.world, sol.S
        N = G * S
        dx = sol.length / N
        s_tot, m_tot = msg_all[:, :, 0].sum(0), msg_all[:, :, 1].sum(0)
        certify(sol, msg_all, st, st["step"] - 1)
        P = msg_all[:sol.rank, :, 0].sum(0)
        left2 = msg_all[(sol.rank - 1) % G][:, 6:8]
        right2 = msg_all[(sol.rank + 1) % G][:, 4:6]
        ext = torch.cat([left2, n_rows.double(), right2], dim=-1) - 1.0          # rho of cells -2 .. S+1
        rho = ext[:, 2:-2]
        C = P[:, None] + torch.cumsum(rho, dim=-1)
        j = (sol.rank * S + torch.arange(S)).double()
        rbar = (s_tot / N)[:, None]
        mu = (0.5 * s_tot - m_tot / N - s_tot / (2 * N))[:, None]
        E = -dx * (C - 0.5 * rho - (j + 0.5) * rbar - mu) + (dx / 24.0) * (ext[:, 3:-1] - ext[:, 1:-3])
        d4 = ext[:, :-4] - 4 * ext[:, 1:-3] + 6 * rho - 4 * ext[:, 3:-1] + ext[:, 4:]
        st["cert"][:, 0] = (d4 ** 2).sum(-1)
        st["cert"][:, 1] = E.abs().amax(-1)
        E_rows.copy_(E.float())

    def certify(sol, msg_all, st, step=None):
        step = st["step"] - 1 if step is None else step
        if step < 0:
            return
        N = sol.world * sol.S
        bound = torch.sqrt(msg_all[:, :, 2].sum(0) / N) * sol.length / (32 * np.sqrt(3))
        if bool((bound > sol.cert_tol * msg_all[:, :, 3].amax(0)).any()):
            st["flag"][0] = min(int(st["flag"][0]), step)

    return {"sums": sums, "field": field, "certify": certify}


def _worker(rank, world, port, nx, radius, steps, field_solve, out, dt=1e-3):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gnn_plasma_flux_b200.domain import (DomainDecomposedBaselineSolver, DomainDecomposedHybridSolver, TorchDistComm,
                                                 split_slabs)
        torch.set_num_threads(2)
        weights = P.init_weights(0)
        grid = P.Grid(nx=nx, dt=dt)
        ics = torch.from_numpy(np.stack([P.stable_initial_condition(grid, s) for s in range(3)]))
        k = torch.as_tensor(grid.k)
        kw = dict(field_fn=lambda n: batched.poisson(n, k), field_solve=field_solve,
                  field_stages=(cpu_field_stages() if field_solve == "alltoall" else
                                cpu_scan_stages() if field_solve == "scan" else None))
        sol = DomainDecomposedHybridSolver(_Model(), nx, dt=dt, graph_radius=radius, rank=rank, world=world,
                                           device="cpu", slab_fn=lambda *a: None, **kw)
        sol._slab_fn = _oracle_slab_fn(weights, grid, radius, sol)
        assert sol.field_mode == field_solve and sol.halo == 4 * radius + 1
        np.testing.assert_array_equal(
            sol.x_ext.numpy(), grid.x[(rank * sol.owned - sol.halo + np.arange(sol.owned + 2 * sol.halo)) % nx].astype(np.float32))
        comm = TorchDistComm()
        local = split_slabs(ics, world)[rank]
        ref = ics
        for _ in range(steps):
            local = sol.step(local, comm)
            ref = batched.hybrid_step(weights, ref, grid.x, grid.k, grid.dt, grid.dx, radius=radius)
        want = split_slabs(ref, world)[rank]
        out[("hybrid", rank)] = float(P.rel_err(local.numpy(), want.numpy()).max())
        if field_solve == "scan":      # collective: every rank calls it.  The network's fluxes are piecewise linear (ReLU),
            # so n' has kinks whose 4th differences use up the conservative bound on a 4096-cell grid within a few
            # steps (at 2^21 cells per rank they are 500 times smaller): reported, not required, here
            verdict = sol.first_uncertified(local, comm)
            assert verdict is None or 0 <= verdict < steps
        # the classical solver on slabs (halo 4, one cell read)
        bs = DomainDecomposedBaselineSolver(nx, dt=dt, nu=1e-3, rank=rank, world=world, device="cpu",
                                            slab_fn=lambda *a: None, **kw)
        bs._slab_fn = _baseline_slab_fn(grid, bs)
        local = split_slabs(ics, world)[rank]
        ref = ics
        for _ in range(steps):
            local = bs.step(local, comm)
            ref = batched.baseline_step(ref, grid.k, grid.dt, grid.dx, grid.nu)
        want = split_slabs(ref, world)[rank]
        out[("baseline", rank)] = float(P.rel_err(local.numpy(), want.numpy()).max())
        if field_solve == "scan":
            assert bs.first_uncertified(local, comm) is None
    finally:
        dist.destroy_process_group()


def _run(field_solve, nx=512, dt=1e-3):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(2, port, nx, 2, 3, field_solve, out, dt), nprocs=2, join=True)
    res = dict(out)
    assert len(res) == 4 and all(v < 2e-6 for v in res.values()), res


def test_two_rank_domain_decomposition_allgather_matches_global_oracle():
    _run("allgather")


def test_two_rank_domain_decomposition_alltoall_matches_global_oracle():
    _run("alltoall")


def test_two_rank_domain_decomposition_scan_matches_global_oracle():
    """The distributed prefix-sum field solve: per-rank sums, one all-gather (gloo), local reconstruction."""
    # 4096 cells: long enough for the trapezoid + Euler-Maclaurin form (512 cells: 5e-6 off, uncertified); dt/dx = 0.02
    # as in the GPU configs (at dt/dx = 0.65 the fp32 noise of the network's fluxes alone exhausts the conservative bound)
    _run("scan", nx=4096, dt=0.02 * 2 * np.pi / 4096)


def test_distributed_scan_solve_stages_cpu():
    """The algebra of the distributed scan solve on 4 virtual ranks: equals the fp64 spectral operator for a smooth
    density, and its certificate flags a rough one."""
    from gnn_plasma_flux_b200.domain import DistributedScanSolve, scan_first_uncertified_emulated, scan_solve_emulated
    nx, world, batch = 4096, 4, 3
    x = (np.arange(nx) + 0.5) * 2 * np.pi / nx
    dens = np.stack([1.1 + 0.2 * np.sin((b + 1) * x + 0.3) + 0.05 * np.cos(3 * x) for b in range(batch)]).astype(np.float32)
    S = nx // world
    grid = P.Grid(nx=nx)
    for rough in (False, True):
        n = torch.from_numpy(dens.copy())
        if rough:
            n[1] += torch.from_numpy((1e-3 * np.random.RandomState(0).randn(nx)).astype(np.float32))
        solvers = [DistributedScanSolve(nx, 2 * np.pi, r, world, "cpu", 1e-5, cpu_scan_stages()) for r in range(world)]
        E = torch.zeros_like(n)
        rows = [n[:, r * S:(r + 1) * S] for r in range(world)]
        scan_solve_emulated(solvers, rows, [E[:, r * S:(r + 1) * S] for r in range(world)])
        want = np.stack([P.solve_poisson(d, grid.k) for d in n.numpy()])
        if not rough:
            assert np.abs(E.numpy() - want).max() <= 3e-7 * np.abs(want).max()
        assert scan_first_uncertified_emulated(solvers, rows) == (0 if rough else None)


def test_distributed_field_solve_stages_cpu():
    """The decomposition's algebra alone, 4 virtual ranks in one process (no process group): the CPU stage stand-ins
    chained by solve_emulated reproduce the fp64 spectral operator, Nyquist and mean components included."""
    from gnn_plasma_flux_b200.domain import DistributedFieldSolve, solve_emulated
    nx, world, batch = 2048, 4, 3
    rng = np.random.RandomState(0)
    x = np.linspace(0, 2 * np.pi, nx, endpoint=False)
    dens = np.stack([1.2 + 0.2 * np.sin((b + 1) * x) + 0.1 * np.cos(np.pi * np.arange(nx)) + 0.02 * rng.randn(nx)
                     for b in range(batch)]).astype(np.float32)
    n = torch.from_numpy(dens)
    S = nx // world
    solvers = [DistributedFieldSolve(nx, 2 * np.pi, r, world, "cpu", cpu_field_stages()) for r in range(world)]
    E = torch.zeros_like(n)
    solve_emulated(solvers, [n[:, r * S:(r + 1) * S] for r in range(world)], [E[:, r * S:(r + 1) * S] for r in range(world)])
    grid = P.Grid(nx=nx)
    want = np.stack([P.solve_poisson(d, grid.k) for d in dens])
    assert np.abs(E.numpy() - want).max() <= 2e-6 * np.abs(want).max()
