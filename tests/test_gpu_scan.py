"""The certified prefix-sum field solve of the classical solver (csrc/scan_poisson.cu) against the oracle.

Reference operator: src/baseline_solver.py:59-68 (E = Re ifft(i fft(n - 1) / k)) and the step :80-101.  The scan solve
must (1) leave the finite-volume update bit-exact, (2) stay within the certified distance of the spectral field, (3)
refuse -- through its certificate -- every input it cannot certify, in which case "auto" repeats with the FFT solve.
"""
import numpy as np
import pytest
import torch

from oracle import batched, ref_port as P

pytestmark = pytest.mark.gpu
STEP_TOL = 1e-5          # north_star: fp32 path within 1e-5 relative per step


def _stable_dt(nx, nu=1e-3):
    dx = 2 * np.pi / nx
    return min(0.02 * dx, 0.2 * dx * dx / nu)


def _ics(nx, B, dt):
    grid = P.Grid(nx=nx, dt=dt, nu=1e-3)
    return grid, np.stack([P.stable_initial_condition(grid, s) for s in range(B)])


def _spectral_fp64(n):
    """The reference operator evaluated in float64 (numpy), [B, nx] -> [B, nx]."""
    nx = n.shape[-1]
    rho = (n - np.float32(1.0)).astype(np.float64)
    k = 2 * np.pi * np.fft.fftfreq(nx, 2 * np.pi / nx)
    rh = np.fft.fft(rho, axis=-1)
    eh = np.zeros_like(rh)
    eh[..., 1:] = 1j * rh[..., 1:] / k[1:]
    return np.real(np.fft.ifft(eh, axis=-1))


@pytest.mark.parametrize("nx,B", [(4096, 2), (12000, 1), (4096 + 8, 3), (1 << 16, 3), (1 << 20, 2)])
def test_scan_rollout_vs_oracle(built_lib, nx, B):
    """n', u' of the first step bit-exact; 5 steps against the CPU port within the per-step tolerance; the field of
    the final state against the float64 spectral operator applied to the final density."""
    from gnn_plasma_flux_b200 import BaselineSolver
    dt = _stable_dt(nx)
    grid, ics = _ics(nx, B, dt)
    sol = BaselineSolver(nx=nx, dt=dt, nu=1e-3, field_solve="scan")
    dev = torch.from_numpy(ics).cuda()
    one, _, flux = sol.rollout(dev, 1, record_flux=True)
    ref1 = batched.baseline_step(torch.from_numpy(ics), grid.k, grid.dt, grid.dx, grid.nu).numpy()
    np.testing.assert_array_equal(one.cpu().numpy()[:, :2], ref1[:, :2])
    np.testing.assert_array_equal(flux[0].cpu().numpy(), ics[:, 0] * ics[:, 1])
    assert P.rel_err(one.cpu().numpy(), ref1).max() <= STEP_TOL          # the fp32 port's own FFT error is ~1.5e-6 at 2^20
    out = sol.rollout(dev, 5)[0].cpu().numpy()
    assert sol.last_field_solve == "scan" and sol.last_uncertified_step is None
    ref = torch.from_numpy(ics)
    for _ in range(5):
        ref = batched.baseline_step(ref, grid.k, grid.dt, grid.dx, grid.nu)
    assert P.rel_err(out, ref.numpy()).max() <= STEP_TOL
    e64 = _spectral_fp64(out[:, 0])
    assert np.abs(out[:, 2] - e64).max() <= 4e-7 * np.abs(e64).max()          # fp32 rounding of E itself
    assert abs(out[:, 2].astype(np.float64).mean()) <= 1e-7 * np.abs(e64).max()


def test_scan_is_closer_to_float64_than_the_fft_solve(built_lib):
    """At 2^22 cells the fp32 FFT solve carries log2(N) rounding steps; the prefix sum accumulates in float64."""
    from gnn_plasma_flux_b200 import BaselineSolver
    nx = 1 << 22
    dt = _stable_dt(nx)
    _, ics = _ics(nx, 1, dt)
    dev = torch.from_numpy(ics).cuda()
    e64 = _spectral_fp64(BaselineSolver(nx=nx, dt=dt, field_solve="scan").rollout(dev, 1)[0][:, 0].cpu().numpy())
    err = {}
    for mode in ("scan", "spectral"):
        out = BaselineSolver(nx=nx, dt=dt, nu=1e-3, field_solve=mode).rollout(dev, 1)[0].cpu().numpy()
        err[mode] = np.abs(out[:, 2] - e64).max() / np.abs(e64).max()
    assert err["scan"] <= 2e-7 and err["scan"] <= err["spectral"], err


def test_scan_certificate_rejects_rough_density(built_lib):
    """White noise on the density: the bound exceeds the tolerance, "scan" raises, "auto" repeats the rollout with the
    FFT solve and returns exactly what "spectral" returns."""
    from gnn_plasma_flux_b200 import BaselineSolver, _lib
    nx = 1 << 16
    dt = _stable_dt(nx)
    _, ics = _ics(nx, 2, dt)
    rng = np.random.RandomState(0)
    ics[:, 0] += (1e-3 * rng.randn(2, nx)).astype(np.float32)
    dev = torch.from_numpy(ics).cuda()
    spectral = BaselineSolver(nx=nx, dt=dt, nu=1e-3, field_solve="spectral").rollout(dev, 4)[0]
    auto = BaselineSolver(nx=nx, dt=dt, nu=1e-3)                                  # default: auto
    got = auto.rollout(dev, 4)[0]
    assert auto.last_field_solve == "spectral" and auto.last_uncertified_step == 1 and auto.last_uncertified_ics == [0, 1]
    assert torch.equal(got, spectral)
    # certificates are per IC: with one smooth and one rough IC only the rough one is repeated
    mixed = dev.clone()
    mixed[0] = torch.from_numpy(_ics(nx, 1, dt)[1][0]).cuda()
    got2, traj2, flux2 = auto.rollout(mixed, 4, record_every=2, record_flux=True)
    assert auto.last_field_solve == "scan+spectral" and auto.last_uncertified_ics == [1]
    want2 = BaselineSolver(nx=nx, dt=dt, nu=1e-3, field_solve="spectral").rollout(mixed, 4, record_every=2, record_flux=True)
    assert torch.equal(got2[1], want2[0][1]) and torch.equal(traj2[:, 1], want2[1][:, 1]) and torch.equal(flux2[:, 1], want2[2][:, 1])
    assert P.rel_err(got2[:1].cpu().numpy(), want2[0][:1].cpu().numpy()).max() <= 2e-6
    with pytest.raises(_lib.FluxGNNError, match="not certified"):
        BaselineSolver(nx=nx, dt=dt, nu=1e-3, field_solve="scan").rollout(dev, 4)
    # the bound is an upper bound: measured deviation of the uncertified scan field from the float64 operator
    loose = BaselineSolver(nx=nx, dt=dt, nu=1e-3, field_solve="scan", cert_tol=1.0)
    out = loose.rollout(dev, 1)[0].cpu().numpy()
    e64 = _spectral_fp64(out[:, 0])
    rho = (out[:, 0] - np.float32(1.0)).astype(np.float64)
    d4 = np.roll(rho, -2, -1) - 4 * np.roll(rho, -1, -1) + 6 * rho - 4 * np.roll(rho, 1, -1) + np.roll(rho, 2, -1)
    bound = np.sqrt((d4 ** 2).mean(-1)) * 2 * np.pi / (32 * np.sqrt(3))
    dev_max = np.abs(out[:, 2] - e64).max(-1)
    assert (dev_max <= bound).all(), (dev_max, bound)         # (very conservative for white noise: Cauchy-Schwarz)


def test_scan_short_or_odd_grids_use_the_fft_solve(built_lib):
    from gnn_plasma_flux_b200 import BaselineSolver, _lib
    for nx in (64, 1024, 4100):
        sol = BaselineSolver(nx=nx, dt=_stable_dt(nx), nu=1e-3)
        _, ics = _ics(nx, 1, sol.dt)
        sol.rollout(torch.from_numpy(ics).cuda(), 2)
        assert sol.last_field_solve == "spectral"
        with pytest.raises(_lib.FluxGNNError, match="nx >= 4096"):
            sol.rollout(torch.from_numpy(ics).cuda(), 2, field_solve="scan")


@pytest.mark.parametrize("record_every", [1, 3])
def test_scan_trajectory_and_flux(built_lib, record_every):
    """Recorded states carry their reconstructed field; fluxes and trajectory agree with the FFT path."""
    from gnn_plasma_flux_b200 import BaselineSolver
    nx, B, steps = 1 << 14, 2, 9
    dt = _stable_dt(nx)
    _, ics = _ics(nx, B, dt)
    dev = torch.from_numpy(ics).cuda()
    a = BaselineSolver(nx=nx, dt=dt, nu=1e-3, field_solve="spectral").rollout(dev, steps, record_every, True)
    b = BaselineSolver(nx=nx, dt=dt, nu=1e-3, field_solve="scan").rollout(dev, steps, record_every, True)
    assert b[1].shape == (steps // record_every, B, 3, nx) and b[2].shape == (steps, B, nx)
    assert torch.equal(b[1][-1], b[0])
    assert torch.equal(a[2][0], b[2][0])                          # F_n of the first step: same inputs
    assert P.rel_err(b[0].cpu().numpy(), a[0].cpu().numpy()).max() <= 5e-6                       # final state
    assert P.rel_err(b[1].cpu().numpy().reshape(-1, 3, nx), a[1].cpu().numpy().reshape(-1, 3, nx)).max() <= 5e-6
    assert float((a[2] - b[2]).abs().max() / a[2].abs().max()) <= 5e-6                            # fluxes
    # run(): the reference's list-of-states API on top of the same call
    states, fluxes = BaselineSolver(nx=nx, dt=dt, nu=1e-3).run(ics[0], n_steps=4)
    assert states.shape == (5, 3, nx) and fluxes.shape == (4, nx)
    np.testing.assert_array_equal(states[0], ics[0])


def test_scan_full_size_c5(built_lib):
    """BASELINE.json configs[4] at full size: 2^24 cells, 10 steps.  Properties: mass conserved to round-off, zero-mean
    field, every field certified, agreement with the FFT path within the fp32 accuracy of the FFT path."""
    from gnn_plasma_flux_b200 import BaselineSolver
    nx = 1 << 24
    dt = _stable_dt(nx)
    _, ics = _ics(nx, 1, dt)
    dev = torch.from_numpy(ics).cuda()
    sol = BaselineSolver(nx=nx, dt=dt, nu=1e-3)
    out = sol.rollout(dev, 10)[0]
    assert sol.last_field_solve == "scan" and sol.last_uncertified_step is None
    ref = sol.rollout(dev, 10, field_solve="spectral")[0]
    assert torch.isfinite(out).all()
    assert torch.equal(out[:, 0], ref[:, 0])                      # the field only enters through dt * E, below rounding here
    assert P.rel_err(out.cpu().numpy(), ref.cpu().numpy()).max() <= 3e-6
    o = out.cpu().numpy()
    assert abs(o[0, 0].astype(np.float64).sum() - ics[0, 0].astype(np.float64).sum()) <= 4 * nx * np.finfo(np.float32).eps
    e64 = _spectral_fp64(o[:, 0])
    assert np.abs(o[:, 2] - e64).max() <= 4e-7 * np.abs(e64).max()


# ----------------------------------------------------------------------------- distributed form (domain decomposition)
@pytest.mark.parametrize("world,nx", [(1, 4096), (4, 1 << 14), (8, 1 << 20), (2, 12000)])
def test_baseline_decomposed_scan_solve(built_lib, world, nx):
    """The classical step on G virtual ranks with the distributed prefix-sum field solve (per-rank sums, one 48-byte
    message per IC and rank, local reconstruction) against the undivided solver with the FFT solve."""
    from gnn_plasma_flux_b200 import BaselineSolver
    from gnn_plasma_flux_b200.domain import (DomainDecomposedBaselineSolver, scan_first_uncertified_emulated, split_slabs,
                                             step_emulated)
    dt = _stable_dt(nx)
    _, ics = _ics(nx, 3, dt)
    dev = torch.from_numpy(ics).cuda()
    whole = BaselineSolver(nx=nx, dt=dt, nu=1e-3, field_solve="spectral")
    solvers = [DomainDecomposedBaselineSolver(nx, dt=dt, nu=1e-3, rank=r, world=world, device="cuda", field_solve="scan")
               for r in range(world)]
    assert solvers[0].field_mode == "scan"
    locals_, ref = split_slabs(dev, world), dev
    for t in range(5):
        locals_ = step_emulated(solvers, locals_)
        ref = whole.rollout(ref, 1)[0]
        got = torch.cat(list(locals_), dim=-1)
        if t == 0:
            assert torch.equal(got[:, :2], ref[:, :2])
    assert P.rel_err(got.cpu().numpy(), ref.cpu().numpy()).max() <= STEP_TOL
    out = got.cpu().numpy()
    e64 = _spectral_fp64(out[:, 0])
    assert np.abs(out[:, 2] - e64).max() <= 4e-7 * np.abs(e64).max()
    assert scan_first_uncertified_emulated([s._scan for s in solvers], [loc[:, 0] for loc in locals_]) is None


def test_decomposed_scan_solve_reports_rough_density(built_lib):
    from gnn_plasma_flux_b200.domain import DomainDecomposedBaselineSolver, scan_first_uncertified_emulated, split_slabs, step_emulated
    nx, world = 1 << 14, 4
    dt = _stable_dt(nx)
    _, ics = _ics(nx, 2, dt)
    ics[1, 0] += (1e-3 * np.random.RandomState(1).randn(nx)).astype(np.float32)       # only the second IC is rough
    solvers = [DomainDecomposedBaselineSolver(nx, dt=dt, nu=1e-3, rank=r, world=world, device="cuda", field_solve="scan")
               for r in range(world)]
    locals_ = split_slabs(torch.from_numpy(ics).cuda(), world)
    for _ in range(3):
        locals_ = step_emulated(solvers, locals_)
    assert scan_first_uncertified_emulated([s._scan for s in solvers], [loc[:, 0] for loc in locals_]) == 0


@pytest.mark.parametrize("precision", ["fp32", "fp16x3"])
@pytest.mark.parametrize("world,nx,radius", [(2, 1 << 12, 3), (4, 1 << 15, 2), (8, 1 << 15, 1), (8, 1 << 20, 3)])
def test_hybrid_decomposed_scan_solve(precision, world, nx, radius):
    """BASELINE.json configs[3] scaled down: the hybrid step on G virtual ranks, halo exchange + distributed prefix-sum
    field solve, against the undivided solver (slabs start at halo = 4r+1 cells: the unaligned load path)."""
    from gnn_plasma_flux_b200 import HybridSolver
    from gnn_plasma_flux_b200.domain import (DomainDecomposedHybridSolver, scan_first_uncertified_emulated, split_slabs,
                                             step_emulated)
    from gnn_plasma_flux_b200.synthetic import seeded_model
    model = seeded_model(0, torch.device("cuda"))
    dt = 0.02 * (2 * np.pi / nx)
    _, ics = _ics(nx, 3, dt)
    dev = torch.from_numpy(ics).cuda()
    whole = HybridSolver(None, radius, nx=nx, dt=dt, device="cuda", graph_radius=radius, model=model, precision=precision)
    solvers = [DomainDecomposedHybridSolver(model, nx, dt=dt, graph_radius=radius, rank=r, world=world, device="cuda",
                                            precision=precision, field_solve="scan") for r in range(world)]
    locals_, ref = split_slabs(dev, world), dev
    for t in range(3):
        locals_ = step_emulated(solvers, locals_)
        ref, _ = whole.rollout(ref, 1)
        got = torch.cat(list(locals_), dim=-1)
        if t == 0:
            assert torch.equal(got[:, :2], ref[:, :2])
            assert float((got[:, 2] - ref[:, 2]).abs().max()) <= 2e-6 * float(ref[:, 2].abs().max())
    assert P.rel_err(got.cpu().numpy(), ref.cpu().numpy()).max() <= STEP_TOL
    # the network's fluxes are piecewise linear (ReLU): n' carries kinks whose 4th differences scale with dx, so the
    # conservative certificate passes on long grids only (BASELINE.json's shape has 2^21 cells per rank)
    verdict = scan_first_uncertified_emulated([s._scan for s in solvers], [loc[:, 0] for loc in locals_])
    assert verdict is None or nx < (1 << 15), verdict


@pytest.mark.parametrize("kind,world,nx,radius", [("baseline", 4, 1 << 14, 0), ("baseline", 1, 4096, 0), ("baseline", 2, 12000, 0),
                                                   ("hybrid", 2, 1 << 12, 3), ("hybrid", 8, 1 << 15, 2)])
def test_peer_memory_step_equals_the_collective_step(built_lib, kind, world, nx, radius):
    """The step over peer memory (halo cells and field-solve messages stored straight into the neighbours' buffers:
    fluxgnn_peer_halo_push, fluxgnn_peer_allgather; G virtual ranks on an EmulatedFabric) against the same solvers'
    collective step (halo exchange + all-gather, emulated): bit-identical n', u', E' over four steps, and n', u' of the
    first step bit-exact against the undivided solver."""
    from gnn_plasma_flux_b200 import BaselineSolver, HybridSolver
    from gnn_plasma_flux_b200.domain import (DomainDecomposedBaselineSolver, DomainDecomposedHybridSolver, EmulatedFabric,
                                             split_slabs, step_emulated, step_peer_emulated)
    from gnn_plasma_flux_b200.synthetic import seeded_model
    if kind == "baseline":
        dt = _stable_dt(nx)
        whole = BaselineSolver(nx=nx, dt=dt, nu=1e-3, field_solve="spectral")
        make = lambda r, fabric: DomainDecomposedBaselineSolver(nx, dt=dt, nu=1e-3, rank=r, world=world, device="cuda",
                                                                field_solve="scan", fabric=fabric)
        first = lambda st: whole.rollout(st, 1)[0]
    else:
        model = seeded_model(0, torch.device("cuda"))
        dt = 0.02 * (2 * np.pi / nx)
        whole = HybridSolver(None, radius, nx=nx, dt=dt, device="cuda", graph_radius=radius, model=model)
        make = lambda r, fabric: DomainDecomposedHybridSolver(model, nx, dt=dt, graph_radius=radius, rank=r, world=world,
                                                              device="cuda", field_solve="scan", fabric=fabric)
        first = lambda st: whole.rollout(st, 1)[0]
    _, ics = _ics(nx, 3, dt)
    dev = torch.from_numpy(ics).cuda()
    fabrics = EmulatedFabric.create(world, "cuda")
    peers = [make(r, fabrics[r]) for r in range(world)]
    colls = [make(r, None) for r in range(world)]
    a, b = split_slabs(dev, world), split_slabs(dev, world)
    for t in range(4):
        a = step_peer_emulated(peers, a)
        b = step_emulated(colls, b)
        got, want = torch.cat(list(a), dim=-1), torch.cat(list(b), dim=-1)
        assert torch.equal(got, want), t
        if t == 0:
            assert torch.equal(got[:, :2], first(dev)[:, :2])
    assert torch.isfinite(got).all()


def test_peer_exchange_entry_points(built_lib):
    """fluxgnn_peer_halo_push and fluxgnn_peer_allgather through the C ABI, three virtual ranks on one GPU: every ghost
    zone receives exactly the neighbour's edge cells of the requested channels, every gather slot the sender's bytes."""
    from gnn_plasma_flux_b200 import _lib
    L = _lib.lib()
    G, B, S, H = 3, 4, 40, 5
    ld = S + 2 * H
    gen = torch.Generator("cuda").manual_seed(3)
    ext = [torch.randn(B, 3, ld, device="cuda", generator=gen) for _ in range(G)]
    before = [e.clone() for e in ext]
    stream = torch.cuda.current_stream().cuda_stream
    for r in range(G):
        _lib.check(L.fluxgnn_peer_halo_push(before[r].data_ptr(), ext[(r - 1) % G].data_ptr(), ext[(r + 1) % G].data_ptr(),
                                            B, S, H, 0, 2, stream), "fluxgnn_peer_halo_push")
    for r in range(G):
        left, right = before[(r - 1) % G], before[(r + 1) % G]
        assert torch.equal(ext[r][:, :2, :H], left[:, :2, S:S + H])             # left ghosts = left neighbour's last H cells
        assert torch.equal(ext[r][:, :2, H + S:], right[:, :2, H:2 * H])        # right ghosts = right neighbour's first H
        assert torch.equal(ext[r][:, 2], before[r][:, 2])                       # channel 2 was not requested
        assert torch.equal(ext[r][:, :, H:H + S], before[r][:, :, H:H + S])     # interiors untouched
    nbytes, offset = 96, 64
    blocks = [torch.zeros(offset + G * nbytes + 32, dtype=torch.uint8, device="cuda") for _ in range(G)]
    bases = torch.tensor([b.data_ptr() for b in blocks], dtype=torch.int64, device="cuda")
    msgs = [torch.randint(0, 255, (nbytes,), dtype=torch.uint8, device="cuda", generator=gen) for _ in range(G)]
    for r in range(G):
        _lib.check(L.fluxgnn_peer_allgather(msgs[r].data_ptr(), nbytes, bases.data_ptr(), offset, r, G, stream),
                   "fluxgnn_peer_allgather")
    for b in blocks:
        assert torch.equal(b[offset:offset + G * nbytes], torch.cat(msgs))
        assert int(b[:offset].sum()) == 0 and int(b[offset + G * nbytes:].sum()) == 0
    assert L.fluxgnn_peer_allgather(msgs[0].data_ptr(), 40, bases.data_ptr(), offset, 0, G, stream) < 0    # not a multiple of 16


@pytest.mark.parametrize("nx,B", [(1 << 20, 16), (1 << 22, 8), (1 << 18, 40)])
def test_scan_repeated_launches_are_bit_stable(built_lib, nx, B):
    """Race hunt (the short form of scripts/stress_scan.py): at large batch every SM holds several CTAs whose bulk-copy
    rings refill while other warps still read; the first step's n', u' must equal the FFT path's bit for bit on every
    repetition (the field is not involved in the first step), and longer rollouts must stay certified."""
    from gnn_plasma_flux_b200 import BaselineSolver
    from gnn_plasma_flux_b200.synthetic import stable_initial_conditions
    dt = 0.2 * (2 * np.pi / nx) ** 2 / 1e-3
    sol = BaselineSolver(nx=nx, dt=dt, nu=1e-3, device="cuda")
    state = stable_initial_conditions(sol, B)
    ref1 = sol.rollout(state, 1, field_solve="spectral")[0]
    for _ in range(12):
        out = sol.rollout(state, 1, field_solve="scan")[0]
        assert torch.equal(out[:, :2], ref1[:, :2])
        sol.rollout(state, 5, field_solve="scan")


def test_scan_more_ics_than_resident_ctas(built_lib):
    """900 ICs of 4096 cells: one segment (CTA) per IC, more CTAs than the device holds at once."""
    from gnn_plasma_flux_b200 import BaselineSolver
    from gnn_plasma_flux_b200.synthetic import stable_initial_conditions
    nx, B = 4096, 900
    # (cert_tol loosened: among 900 random ICs a few have fields of ~3e-3, for which the rounding-noise part of the
    #  conservative bound, ~3e-8 and growing with the square root of the step count, exceeds 1e-5 * max|E| -- "auto"
    #  would repeat those with the FFT solve; the subject here is the launch geometry)
    sol = BaselineSolver(nx=nx, dt=_stable_dt(nx), nu=1e-3, device="cuda", cert_tol=1e-3)
    state = stable_initial_conditions(sol, B)
    ref = sol.rollout(state, 4, field_solve="spectral")[0]
    got = sol.rollout(state, 4, field_solve="scan")[0]
    assert sol.last_field_solve == "scan"
    assert P.rel_err(got.cpu().numpy(), ref.cpu().numpy()).max() <= 2e-6
    one = sol.rollout(state, 1, field_solve="scan")[0]
    assert torch.equal(one[:, :2], sol.rollout(state, 1, field_solve="spectral")[0][:, :2])


def test_auto_rollout_falls_back_chunk_by_chunk(built_lib, monkeypatch):
    """The rounding noise of n -- and with it the conservative bound -- can only grow with the step count, so a long
    "auto" rollout runs in chunks: the first chunk whose certificate fails, and all later ones, are repeated with the FFT
    solve.  The failure is injected here (third chunk reports step 5 of the chunk as uncertified)."""
    from gnn_plasma_flux_b200 import BaselineSolver
    monkeypatch.setattr(BaselineSolver, "AUTO_CHUNK", 16)
    nx, B, steps = 1 << 16, 2, 100
    dt = _stable_dt(nx)
    _, ics = _ics(nx, B, dt)
    dev = torch.from_numpy(ics).cuda()
    sol = BaselineSolver(nx=nx, dt=dt, nu=1e-3)
    real, calls = sol._rollout_scan, []

    def flaky(state, n_steps, record_every, traj, flux):
        out, bad = real(state, n_steps, record_every, traj, flux)       # bad: int32[B], INT_MAX = certified
        calls.append(n_steps)
        if len(calls) == 3:
            bad = bad.clone()
            bad[1] = 5                                                 # IC 1 reports step 5 of this chunk as uncertified
        return out, bad

    monkeypatch.setattr(sol, "_rollout_scan", flaky)
    got = sol.rollout(dev, steps, record_every=4, record_flux=True)
    assert calls == [16] * 6 + [4]                                # IC 0 stays on the scan path to the end (100 = 6 x 16 + 4)
    assert sol.last_field_solve == "scan+spectral" and sol.last_uncertified_step == 32 + 5 and sol.last_uncertified_ics == [1]
    ref = BaselineSolver(nx=nx, dt=dt, nu=1e-3, field_solve="spectral").rollout(dev, steps, record_every=4, record_flux=True)
    assert got[1].shape == ref[1].shape == (steps // 4, B, 3, nx) and got[2].shape == (steps, B, nx)
    assert torch.equal(got[1][-1], got[0])
    assert P.rel_err(got[0].cpu().numpy(), ref[0].cpu().numpy()).max() <= 5e-6
    assert P.rel_err(got[1].cpu().numpy().reshape(-1, 3, nx), ref[1].cpu().numpy().reshape(-1, 3, nx)).max() <= 5e-6
    assert float((got[2] - ref[2]).abs().max() / ref[2].abs().max()) <= 5e-6
    # without the injected failure the same rollout is certified throughout, chunk after chunk
    easy = BaselineSolver(nx=nx, dt=dt, nu=1e-3)
    out = easy.rollout(dev, steps, record_every=4)
    assert easy.last_field_solve == "scan" and easy.last_uncertified_step is None
    one_call = BaselineSolver(nx=nx, dt=dt, nu=1e-3, field_solve="scan").rollout(dev, steps, record_every=4)
    assert P.rel_err(out[0].cpu().numpy(), one_call[0].cpu().numpy()).max() <= 1e-6     # chunks re-materialise E: same to rounding
