"""CPU suite: the oracle against the golden vectors frozen from the live reference
(oracle/make_golden.py), and the internal consistency of its two restatements."""
import numpy as np
import torch

from conftest import load_golden
from oracle import batched, ref_port as P


def test_weights_match_seeded_init(weights):
    w = P.init_weights(0)
    assert set(w) == set(weights)
    for key in w:
        np.testing.assert_array_equal(w[key], weights[key])
    assert sum(v.size for v in w.values()) == 165249


def test_ring_edges_reference_prefix():
    for nx in (4, 64):
        e1 = P.ring_edges(nx, 1)
        assert e1.shape == (2, 2 * nx) and e1.dtype == np.int64
        src = np.arange(nx)
        np.testing.assert_array_equal(e1[0, :nx], src)
        np.testing.assert_array_equal(e1[1, :nx], (src + 1) % nx)
        np.testing.assert_array_equal(e1[0, nx:], (src + 1) % nx)
        np.testing.assert_array_equal(e1[1, nx:], src)
        e3 = P.ring_edges(nx, 3)
        assert e3.shape == (2, 6 * nx)
        np.testing.assert_array_equal(e3[:, :2 * nx], e1)


def test_forward_golden(weights):
    g1 = load_golden("g1_forward.npz")
    for nx in (64, 1024):
        grid = P.Grid(nx=nx)
        feats = P.node_features(g1[f"state_nx{nx}"], grid.x)
        for r in (1, 2, 3):
            out = P.fluxgnn_forward(weights, feats, P.ring_edges(nx, r))
            ref = g1[f"flux_nx{nx}_r{r}"]
            assert out.shape == ref.shape == (2 * r * nx,)
            # same primitives, same order; BLAS blocking may differ between thread counts
            assert np.abs(out - ref).max() <= 2e-6 * np.abs(ref).max()


def test_poisson_and_ic_golden():
    g4 = load_golden("g4_poisson_ic.npz")
    for nx in (64, 96, 1024):
        grid = P.Grid(nx=nx)
        for seed in (0, 1, 123):
            np.testing.assert_array_equal(P.initial_condition(grid, seed), g4[f"ic_nx{nx}_s{seed}"])
        for name in ("modes", "white", "nyquist", "const"):
            E = P.solve_poisson(g4[f"n_{name}_nx{nx}"], grid.k)
            np.testing.assert_array_equal(E, g4[f"E_{name}_nx{nx}"])
        assert np.abs(g4[f"E_nyquist_nx{nx}"]).max() < 1e-7      # Nyquist annihilated
        assert np.abs(g4[f"E_const_nx{nx}"]).max() == 0.0        # k = 0 removed
        # the circular-convolution form of the operator
        g = P.poisson_kernel(nx, grid.length)
        rho = (g4[f"n_white_nx{nx}"] - np.float32(1.0)).astype(np.float64)
        conv = np.array([np.dot(g[(j - np.arange(nx)) % nx], rho) for j in range(nx)])
        ref = g4[f"E_white_nx{nx}"]
        assert np.abs(conv - ref).max() <= 2e-6 * np.abs(ref).max()


def test_hybrid_step_and_c1_rollout_golden(weights):
    g = load_golden("g23_hybrid_c1.npz")
    grid = P.Grid(nx=64, dt=5e-3)
    for s in (0, 5, 19):
        out = P.hybrid_step(weights, g["ics"][s], grid)
        assert P.rel_err(out, g["step1"][s]).max() < 2e-6
    roll = P.hybrid_run(weights, g["ics"][2], grid, 30)
    assert roll.shape == (31, 3, 64) and roll.dtype == np.float32
    assert P.rel_err(roll[-1], g["rollout"][2][-1]).max() < 1e-4
    np.testing.assert_array_equal(roll[0], g["ics"][2])


def test_batched_matches_port(weights):
    g = load_golden("g23_hybrid_c1.npz")
    grid = P.Grid(nx=64, dt=5e-3)
    out = batched.hybrid_step(weights, torch.from_numpy(g["ics"]), grid.x, grid.k, grid.dt, grid.dx).numpy()
    assert P.rel_err(out, g["step1"]).max() < 2e-6
    gr = load_golden("g2_hybrid_radius.npz")
    g1k = P.Grid(nx=1024, dt=3e-4)
    for r in (1, 2, 3):
        out = batched.hybrid_step(weights, torch.from_numpy(gr["ic_nx1024"]), g1k.x, g1k.k, g1k.dt, g1k.dx, radius=r).numpy()
        assert P.rel_err(out, gr[f"step_nx1024_r{r}"]).max() < 5e-6
    # tiny ring where hops wrap onto themselves
    g1 = load_golden("g1_forward.npz")
    g4x = P.Grid(nx=4)
    fl = batched.edge_fluxes(weights, torch.from_numpy(g1["state_nx4"])[None],
                             torch.from_numpy(g4x.x.astype(np.float32)), 3)[0].numpy()
    assert np.abs(fl - g1["flux_nx4_r3"]).max() <= 2e-6 * np.abs(g1["flux_nx4_r3"]).max()


def test_baseline_golden():
    g5 = load_golden("g5_baseline.npz")
    for nx in (64, 1024):
        grid = P.Grid(nx=nx, dt=float(g5[f"dt_nx{nx}"]), nu=1e-3)
        states, fluxes = g5[f"states_nx{nx}"], g5[f"fluxes_nx{nx}"]
        ps, pf = P.baseline_run(states[0], grid, n_steps=len(fluxes))
        np.testing.assert_array_equal(ps, states)
        np.testing.assert_array_equal(pf, fluxes)
        bt = batched.baseline_step(torch.from_numpy(states[:1]), grid.k, grid.dt, grid.dx, grid.nu).numpy()
        np.testing.assert_array_equal(bt[0, :2], states[1, :2])        # n', u' bit-exact
        assert P.rel_err(bt[0], states[1]).max() < 2e-6


def test_long_rollout_noise_floor_golden():
    g6 = load_golden("g6_long_rollout.npz")
    floor = P.rel_err(g6["ref_fp32"][:, -1], g6["fp64"][:, -1])
    assert np.isfinite(g6["ref_fp32"]).all()
    assert floor.max() < 1e-3 and floor.min() > 1e-7           # the reference's own fp32 noise
    # mass conservation of the reference to round-off
    mass0 = g6["ref_fp32"][:, 0, 0].astype(np.float64).sum(-1)
    massT = g6["ref_fp32"][:, -1, 0].astype(np.float64).sum(-1)
    assert np.abs(massT - mass0).max() < 1000 * 64 * np.finfo(np.float32).eps


def test_long_rollout_at_config_radius_golden(weights):
    """g6b: the port walks the reference-object trajectory bit for bit at the radii of C2 / C3
    (first 100 steps of one IC each), and the frozen fp32 trajectories stay finite and conserve mass."""
    torch.set_num_threads(1)
    g = load_golden("g6b_long_rollout_radius.npz")
    for tag in ("c2", "c3"):
        nx, dt, r, every = int(g[f"{tag}_nx"]), float(g[f"{tag}_dt"]), int(g[f"{tag}_radius"]), int(g[f"{tag}_every"])
        grid = P.Grid(nx=nx, dt=dt)
        ic = g[f"{tag}_ics"][0]
        np.testing.assert_array_equal(P.stable_initial_condition(grid, 50), ic)
        steps = every if tag == "c2" else 10
        run = P.hybrid_run(weights, ic, grid, steps, radius=r)
        np.testing.assert_array_equal(run[0], g[f"{tag}_ref_fp32"][0, 0])
        if tag == "c2":
            np.testing.assert_array_equal(run[every], g[f"{tag}_ref_fp32"][0, 1])
        ref32 = g[f"{tag}_ref_fp32"]
        assert np.isfinite(ref32).all()
        T = int(g[f"{tag}_steps"])
        mass0 = ref32[:, 0, 0].astype(np.float64).sum(-1)
        assert np.abs(ref32[:, -1, 0].astype(np.float64).sum(-1) - mass0).max() < T * nx * np.finfo(np.float32).eps


def test_generic_architecture_golden():
    """G10: the port at architectures other than MODEL_CONFIG's against the reference model's frozen outputs."""
    g = load_golden("g10_generic_architectures.npz")
    for tag, (F, H, L, seed) in {"f4h64l3": (4, 64, 3, 5), "f2h32l2": (2, 32, 2, 6), "f4h16l1": (4, 16, 1, 7)}.items():
        w = P.init_weights(seed, F, H, L)
        for nx in (64, 300):
            for r in (1, 3):
                out = P.fluxgnn_forward(w, g[f"{tag}_feats_nx{nx}"], P.ring_edges(nx, r))
                np.testing.assert_allclose(out, g[f"{tag}_flux_nx{nx}_r{r}"], rtol=0, atol=2e-6 * np.abs(out).max())
        if F == 4:
            run = P.hybrid_run(w, g[f"{tag}_ics"][0], P.Grid(nx=64, dt=5e-3), 5, radius=2)
            assert P.rel_err(run[-1], g[f"{tag}_rollout5_r2"][0, -1]).max() <= 2e-6
            closed = batched.hybrid_run(w, torch.from_numpy(g[f"{tag}_ics"]), P.Grid(nx=64).x, P.Grid(nx=64).k, 5e-3,
                                        P.Grid(nx=64).dx, 5, radius=2).numpy()
            assert P.rel_err(closed, g[f"{tag}_rollout5_r2"][:, -1]).max() <= 1e-5


def test_metrics_and_datagen_golden():
    """the oracle's restatements of evaluate_all.compute_metrics and
    generate_data.generate_dataset against outputs of the reference's own functions."""
    g7 = load_golden("g7_metrics_datagen.npz")
    g23 = load_golden("g23_hybrid_c1.npz")
    m = P.compute_metrics(g23["rollout"][0], g7["truth"])
    for key, val in m.items():
        np.testing.assert_allclose(np.asarray(val, dtype=np.float64), g7["metric_" + key], rtol=1e-12)
    st, fl, nxt = P.generate_dataset(nx=64, num_initial_conditions=3, steps_per_ic=5)
    np.testing.assert_array_equal(st, g7["ds_state_t"])
    np.testing.assert_array_equal(fl, g7["ds_flux_t"])
    np.testing.assert_array_equal(nxt, g7["ds_state_next"])


def test_comparison_models_golden():
    """the PureGNN / PINN restatements against the reference classes' frozen outputs."""
    g = load_golden("g9_comparison_models.npz")
    grid = P.Grid(nx=64)
    x32 = grid.x.astype(np.float32)
    for hidden, layers in ((64, 3), (128, 4)):
        w = P.init_pure_gnn_weights(7, 4, hidden, layers)
        feats = P.node_features(g["ics"][0], grid.x)
        np.testing.assert_array_equal(P.pure_gnn_forward(w, feats, P.ring_edges(64, 1)), g[f"pgnn{hidden}_delta"])
        np.testing.assert_array_equal(P.pure_gnn_rollout(w, g["ics"][1], x32, 10), g[f"pgnn{hidden}_rollout10"][1])
    wp = P.init_pinn_weights(11, 192, 256, 4)
    np.testing.assert_array_equal(P.pinn_forward(wp, g["ics"]), g["pinn_step"])


def test_training_rollout_gradients_golden(weights):
    """Golden g11 (the reference's multi-step training rollout under its OWN autograd, train_ablation.py:172-206) against
    fp64 autograd of the oracle's restatement batched.training_rollout: the same loss, the same gradients w.r.t. every
    parameter and the initial state.  This is the checker HybridSolver.rollout_with_grad is held to on the GPU."""
    g = load_golden("g11_training_rollout.npz")
    for tag, nx, radius, dt, steps in (("nx64_r1", 64, 1, 5e-3, 3), ("nx64_r3", 64, 3, 5e-3, 3), ("nx256_r2", 256, 2, 1e-3, 2)):
        grid = P.Grid(nx=nx, dt=dt)
        wt = {k: torch.from_numpy(v).double().requires_grad_(True) for k, v in weights.items()}
        st0 = torch.from_numpy(g[f"{tag}_ics"]).double().requires_grad_(True)
        states, faces = batched.training_rollout(wt, st0, grid.x, grid.k, grid.dt, grid.dx, steps, radius=radius)
        energies = torch.stack([0.5 * torch.mean(s[:, 1] ** 2, dim=1) for s in states[:-1]])
        loss = torch.mean((energies - energies[0]) ** 2, dim=0).sum()
        loss = loss + (torch.stack(faces, dim=1) * torch.from_numpy(g[f"{tag}_cot_face"]).double()).sum()
        loss = loss + (states[-1][:, 0] * torch.from_numpy(g[f"{tag}_cot_n"]).double()).sum()
        loss = loss + (states[-1][:, 1] * torch.from_numpy(g[f"{tag}_cot_u"]).double()).sum()
        loss.backward()
        rel = lambda a, b: np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)).max() / max(np.abs(b).max(), 1e-12)
        assert P.rel_err(states[-1].detach().numpy().astype(np.float32), g[f"{tag}_final"]).max() <= 3e-5, tag
        assert abs(loss.item() - float(g[f"{tag}_loss"])) <= 1e-4 * abs(float(g[f"{tag}_loss"])), tag
        assert rel(st0.grad.numpy(), g[f"{tag}_dstate0"]) <= 2e-4, tag
        for name, p_ in wt.items():
            gr = p_.grad.numpy()
            if f"{tag}_grad_{name}" in g:
                assert rel(gr, g[f"{tag}_grad_{name}"]) <= 2e-4, (tag, name)
            else:
                assert abs(np.linalg.norm(gr) / float(g[f"{tag}_gradnorm_{name}"]) - 1) <= 2e-4, (tag, name)
                scale = np.abs(gr).max()
                assert np.abs(gr[:8, :8] - g[f"{tag}_gradcorner_{name}"]).max() <= 2e-4 * scale, (tag, name)
                assert np.abs(gr[[5, 77], :] - g[f"{tag}_gradrows_{name}"]).max() <= 2e-4 * scale, (tag, name)
