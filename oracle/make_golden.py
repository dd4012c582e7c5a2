"""Pin the oracle to the live reference and freeze golden vectors.

Runs ONLY in the build container, where the unmodified reference is mounted at
/root/reference (it does not exist on the GPU box).  It

  1. imports the reference's own FluxGNN / build_chain_graph / HybridSolver /
     BaselineSolver,
  2. checks every function of oracle/ref_port.py against them on seeded inputs
     (bit-exact in fp32: same primitives, same order, same machine), and
  3. writes the REFERENCE's outputs as fixtures under tests/golden/.

    python -m oracle.make_golden            # from the repo root

Environment recorded in tests/golden/MANIFEST.json (numpy >= 2 keeps the
forward FFT of a float32 density in complex64; the reference's pinned numpy
1.26.4 would use complex128 -- a ~1e-7 relative effect on E).
"""
from __future__ import annotations

import json
import os
import sys
import tempfile

import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "src"))      # `from config import MODEL_CONFIG` (src/hybrid_solver.py:21)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from src.baseline_solver import BaselineSolver            # noqa: E402  (reference)
from src.flux_gnn import FluxGNN                          # noqa: E402  (reference)
from src.graph_constructor import build_chain_graph       # noqa: E402  (reference)
from src.hybrid_solver import HybridSolver                # noqa: E402  (reference)

from oracle import batched, ref_port as P                 # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def same(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype, (what, a.shape, b.shape, a.dtype, b.dtype)
    if not np.array_equal(a, b):
        raise AssertionError(f"{what}: port differs from reference, max |d| = {np.abs(a - b).max()}")
    print(f"  ok  {what}  {a.shape} {a.dtype}")


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)          # one thread: BLAS blocking is deterministic
    torch.manual_seed(0)
    model = FluxGNN(input_dim=4, hidden_dim=128, num_layers=4).eval()
    weights = {k: v.numpy().copy() for k, v in model.state_dict().items()}
    same_w = P.init_weights(0)
    for key in weights:
        same(same_w[key], weights[key], f"init_weights[{key}]")
    np.savez_compressed(os.path.join(OUT, "weights_seed0.npz"), **weights)
    ckpt = os.path.join(tempfile.mkdtemp(), "w.pt")
    torch.save(model.state_dict(), ckpt)

    # ---- G4 field solve + grid + initial conditions ----------------------
    g4 = {}
    for nx in (64, 1024, 96):
        ref = BaselineSolver(nx=nx)
        g = P.Grid(nx=nx)
        same(g.x, ref.x, f"grid.x nx={nx}")
        same(g.k, ref.k, f"grid.k nx={nx}")
        for seed in (0, 1, 123):
            ic = ref.initial_condition(seed=seed)
            same(P.initial_condition(g, seed), ic, f"initial_condition nx={nx} seed={seed}")
            g4[f"ic_nx{nx}_s{seed}"] = ic
        rng = np.random.RandomState(7)
        cases = {
            "modes": ref.initial_condition(seed=5)[0],
            "white": (1.0 + 0.3 * rng.randn(nx)).astype(np.float32),
            "nyquist": (1.0 + 0.2 * np.cos(np.pi * np.arange(nx))).astype(np.float32),
            "const": np.full(nx, 1.7, dtype=np.float32),
        }
        for name, dens in cases.items():
            e_ref = ref.solve_poisson(dens)
            same(P.solve_poisson(dens, g.k), e_ref, f"solve_poisson {name} nx={nx}")
            conv = np.real(np.fft.ifft(np.fft.fft(P.poisson_kernel(nx, g.length)) * np.fft.fft((dens - 1.0).astype(np.float64))))
            assert np.abs(conv - e_ref).max() <= 2e-6 * max(1.0, np.abs(e_ref).max()), name
            g4[f"n_{name}_nx{nx}"] = dens
            g4[f"E_{name}_nx{nx}"] = e_ref
    np.savez_compressed(os.path.join(OUT, "g4_poisson_ic.npz"), **g4)

    # ---- G1 FluxGNN.forward on ring graphs --------------------------------
    g1 = {}
    for nx in (64, 1024):
        ref = BaselineSolver(nx=nx)
        g = P.Grid(nx=nx)
        state = ref.initial_condition(seed=11)
        nf, ei = build_chain_graph(state, ref.x)
        same(P.node_features(state, g.x), nf.numpy(), f"node_features nx={nx}")
        same(P.ring_edges(nx, 1), ei.numpy(), f"ring_edges r=1 nx={nx}")
        g1[f"state_nx{nx}"] = state
        for r in (1, 2, 3):
            edges = P.ring_edges(nx, r)
            assert np.array_equal(edges[:, :2 * nx], ei.numpy())
            with torch.no_grad():
                out = model(nf, torch.from_numpy(edges)).numpy()
            same(P.fluxgnn_forward(weights, nf.numpy(), edges), out, f"forward nx={nx} r={r}")
            closed = batched.edge_fluxes(weights, torch.from_numpy(state)[None], torch.from_numpy(g.x.astype(np.float32)), r)[0].numpy()
            err = np.abs(closed - out).max() / np.abs(out).max()
            assert err < 2e-6, (nx, r, err)
            print(f"      closed form vs reference forward: rel {err:.2e}")
            g1[f"flux_nx{nx}_r{r}"] = out
    # a tiny ring where hops wrap onto themselves (nx <= 2r)
    state = np.random.RandomState(3).randn(3, 4).astype(np.float32)
    g = P.Grid(nx=4)
    nf = torch.from_numpy(P.node_features(state, g.x))
    with torch.no_grad():
        out = model(nf, torch.from_numpy(P.ring_edges(4, 3))).numpy()
    g1["state_nx4"], g1["flux_nx4_r3"] = state, out
    np.savez_compressed(os.path.join(OUT, "g1_forward.npz"), **g1)

    # ---- G2 / G3 hybrid step and the repo-default rollout (C1) ------------
    hs = HybridSolver(ckpt, 1, nx=64, dt=5e-3, device="cpu")
    g = P.Grid(nx=64, dt=5e-3)
    ics = np.stack([hs.baseline.initial_condition(seed=s) for s in range(20)])
    step1 = np.stack([hs.step(ic) for ic in ics])
    for s in (0, 7, 19):
        same(P.hybrid_step(weights, ics[s], g), step1[s], f"hybrid_step seed={s}")
    roll = np.stack([hs.run(ic, n_steps=30) for ic in ics])           # [20,31,3,64]
    same(P.hybrid_run(weights, ics[3], g, 30), roll[3], "hybrid_run 30 steps seed=3")
    bt = batched.hybrid_step(weights, torch.from_numpy(ics), g.x, g.k, g.dt, g.dx).numpy()
    print("      batched closed form vs reference step: rel", P.rel_err(bt, step1))
    assert P.rel_err(bt, step1).max() < 2e-6
    np.savez_compressed(os.path.join(OUT, "g23_hybrid_c1.npz"), ics=ics, step1=step1, rollout=roll,
                        dt=5e-3, nx=64, radius=1)

    # hybrid step at nx=1024 (reference semantics, radius 1) and radius-r steps
    # (reference forward on the r-ring + reference FV arithmetic, via the port)
    hs1k = HybridSolver(ckpt, 1, nx=1024, dt=3e-4, device="cpu")
    g1k = P.Grid(nx=1024, dt=3e-4)
    ic1k = np.stack([P.stable_initial_condition(g1k, s) for s in range(2)])
    st1k = np.stack([hs1k.step(ic) for ic in ic1k])
    same(P.hybrid_step(weights, ic1k[0], g1k), st1k[0], "hybrid_step nx=1024")
    extra = {"ic_nx1024": ic1k, "step_nx1024_r1": st1k}
    for r in (2, 3):
        extra[f"step_nx1024_r{r}"] = np.stack([P.hybrid_step(weights, ic, g1k, radius=r) for ic in ic1k])
        extra[f"step_nx64_r{r}"] = np.stack([P.hybrid_step(weights, ic, g, radius=r) for ic in ics[:4]])
    np.savez_compressed(os.path.join(OUT, "g2_hybrid_radius.npz"), **extra)

    # ---- G5 baseline solver ------------------------------------------------
    g5 = {}
    for nx, dt, steps in ((64, 5e-3, 40), (1024, 3e-4, 10)):
        ref = BaselineSolver(nx=nx, dt=dt, nu=1e-3)
        g = P.Grid(nx=nx, dt=dt, nu=1e-3)
        ic = ref.initial_condition(seed=2)
        new, fn = ref.step(ic, return_flux=True)
        pn, pf = P.baseline_step(ic, g, return_flux=True)
        same(pn, new, f"baseline_step nx={nx}")
        same(pf, fn, f"baseline_step flux nx={nx}")
        states, fluxes = ref.run(ic, n_steps=steps)
        ps, pfl = P.baseline_run(ic, g, steps)
        same(ps, states, f"baseline_run nx={nx}")
        same(pfl, fluxes, f"baseline_run fluxes nx={nx}")
        g5[f"states_nx{nx}"], g5[f"fluxes_nx{nx}"], g5[f"dt_nx{nx}"] = states, fluxes, dt
    np.savez_compressed(os.path.join(OUT, "g5_baseline.npz"), **g5)

    # ---- G6 1000-step stabilised rollout + fp64 noise floor -----------------
    hs = HybridSolver(ckpt, 1, nx=64, dt=1e-3, device="cpu")
    g = P.Grid(nx=64, dt=1e-3)
    g6_ic = np.stack([P.stable_initial_condition(g, s) for s in range(4)])
    snaps32, snaps64 = [], []
    for ic in g6_ic:
        r32 = hs.run(ic, n_steps=1000)
        assert np.isfinite(r32).all()
        r64 = P.hybrid_run(weights, ic, g, 1000, dtype=torch.float64)
        snaps32.append(r32[::100])
        snaps64.append(r64[::100])
    snaps32, snaps64 = np.stack(snaps32), np.stack(snaps64)       # [4,11,3,64]
    floor = P.rel_err(snaps32[:, -1], snaps64[:, -1])
    print("      reference fp32 vs fp64 after 1000 steps (n,u,E):", floor)
    np.savez_compressed(os.path.join(OUT, "g6_long_rollout.npz"), ics=g6_ic, ref_fp32=snaps32,
                        fp64=snaps64, dt=1e-3, nx=64, radius=1, every=100)

    # ---- G7 evaluation metrics and data generation --------------------
    sys.path.insert(0, os.path.join(REF, "scripts", "training"))      # evaluate_all imports train_pure_gnn bare
    from scripts.evaluation.evaluate_all import compute_metrics as ref_metrics      # reference
    from scripts.training.generate_data import generate_dataset as ref_generate     # reference
    g23 = dict(np.load(os.path.join(OUT, "g23_hybrid_c1.npz")))
    pred = g23["rollout"][0]                                              # hybrid trajectory [31,3,64]
    base = BaselineSolver(nx=64, dt=5e-3)
    truth, _ = base.run(g23["ics"][0], n_steps=30)
    m_ref, m_port = ref_metrics(pred, truth), P.compute_metrics(pred, truth)
    g7 = {"truth": truth}
    for key, val in m_ref.items():
        np.testing.assert_allclose(np.asarray(m_port[key], dtype=np.float64), np.asarray(val, dtype=np.float64), rtol=1e-12)
        g7["metric_" + key] = np.asarray(val, dtype=np.float64)
    print("  ok  compute_metrics vs reference")
    tmp_npz = os.path.join(tempfile.mkdtemp(), "d.npz")
    st_, fl_, nx_, x_, dt_, dx_, nu_ = ref_generate(nx=64, num_initial_conditions=3, steps_per_ic=5, out_path=tmp_npz)
    ps, pf, pn = P.generate_dataset(nx=64, num_initial_conditions=3, steps_per_ic=5)
    same(ps, st_, "generate_dataset state_t")
    same(pf, fl_, "generate_dataset flux_t")
    same(pn, nx_, "generate_dataset state_next")
    g7.update(ds_state_t=st_, ds_flux_t=fl_, ds_state_next=nx_, ds_x=x_, ds_dx=dx_)
    np.savez_compressed(os.path.join(OUT, "g7_metrics_datagen.npz"), **g7)

    # ---- G8 gradients of the reference model through autograd ------------------
    g8 = {}
    state = BaselineSolver(nx=64).initial_condition(seed=21)
    for r in (1, 2):
        ref_model = FluxGNN(input_dim=4, hidden_dim=128, num_layers=4)
        ref_model.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()})
        nf = torch.from_numpy(P.node_features(state, P.Grid(nx=64).x)).requires_grad_(True)
        edges = torch.from_numpy(P.ring_edges(64, r))
        cot = torch.from_numpy(np.random.RandomState(100 + r).randn(2 * r * 64).astype(np.float32))
        (ref_model(nf, edges) * cot).sum().backward()
        g8[f"cot_r{r}"] = cot.numpy()
        g8[f"dfeat_r{r}"] = nf.grad.numpy()
        for name, p_ in ref_model.named_parameters():
            g = p_.grad.numpy()
            if g.size <= 1024:
                g8[f"grad_r{r}_{name}"] = g
            else:                                            # big matrices: norm + a corner + a few rows
                g8[f"gradnorm_r{r}_{name}"] = np.float64(np.linalg.norm(g.astype(np.float64)))
                g8[f"gradcorner_r{r}_{name}"] = g[:8, :8].copy()
                g8[f"gradrows_r{r}_{name}"] = g[[5, 77], :].copy()
        # the oracle's closed form under fp64 autograd agrees with the reference's autograd
        wt = {k: torch.from_numpy(v).double().requires_grad_(True) for k, v in weights.items()}
        st = torch.from_numpy(state).double()[None].requires_grad_(True)
        fl = batched.edge_fluxes(wt, st, torch.from_numpy(P.Grid(nx=64).x.astype(np.float32)), r, hops=r)[0]
        (fl * cot.double()).sum().backward()
        for name, p_ in ref_model.named_parameters():
            ref_g, orc_g = p_.grad.double().numpy(), wt[name].grad.numpy()
            assert np.abs(ref_g - orc_g).max() <= 2e-4 * max(np.abs(ref_g).max(), 1e-6), (r, name)
        assert np.abs(st.grad[0].numpy().T - nf.grad.numpy()[:, :3]).max() <= 2e-4 * np.abs(nf.grad.numpy()).max()
        print(f"  ok  reference autograd vs oracle fp64 autograd, radius {r}")
    g8["state"] = state
    np.savez_compressed(os.path.join(OUT, "g8_gradients.npz"), **g8)

    comparison_goldens()
    long_rollout_goldens()
    generic_architecture_goldens()
    training_rollout_goldens()

    manifest = {
        "generated_by": "oracle/make_golden.py",
        "reference": "/root/reference (shanedirksen/gnn-plasma-flux, unmodified)",
        "numpy": np.__version__, "torch": torch.__version__, "torch_threads": 1,
        "forward_fft": "complex64 (numpy>=2 on float32 input)",
        "noise_floor_1000_steps_fp32_vs_fp64": floor.tolist(),
        "files": sorted(f for f in os.listdir(OUT) if f.endswith(".npz")),
    }
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as fh:
        json.dump(manifest, fh, indent=1)
    print("goldens written to", OUT)


def reference_radius_run(ref_model, ref_base, state0, n_steps, radius, every):
    """The reference's HybridSolver.step/run arithmetic (src/hybrid_solver.py:34-73) driven with the
    REFERENCE objects -- its FluxGNN.forward fed the radius-r ring, its BaselineSolver.solve_poisson --
    for graph radii the reference's own step cannot take (it slices `flux_edge[nx:]`, :46, which only has
    the right length on the nearest-neighbour ring).  Only hop 1 enters the face flux."""
    nx = ref_base.nx
    edges = torch.from_numpy(P.ring_edges(nx, radius))
    c = ref_base.dt / ref_base.dx
    state = state0.astype(np.float32)
    snaps = [state]
    for t in range(1, n_steps + 1):
        n, u, E = state
        nf, _ = build_chain_graph(state, ref_base.x)
        with torch.no_grad():
            flux = ref_model(nf, edges).cpu().numpy()
        face = 0.5 * (flux[:nx] + flux[nx:2 * nx]).astype(np.float32)
        n_new = n - c * (face - np.roll(face, 1))
        fu = 0.5 * u * u
        u_new = (u - c * (fu - np.roll(fu, 1))) + ref_base.dt * E
        state = np.stack([n_new, u_new, ref_base.solve_poisson(n_new)], axis=0).astype(np.float32)
        if t % every == 0:
            snaps.append(state)
    return np.stack(snaps, axis=0)


def long_rollout_goldens():
    """G6b: the long-rollout gate at the radii BASELINE.json's configs really use -- C2 (nx=64, radius 3,
    1000 steps) and C3 (nx=1024, radius 2, 300 steps): reference objects in fp32 + the fp64 restatement.
    `python -m oracle.make_golden --only-g6b` regenerates just this file."""
    torch.set_num_threads(1)
    torch.manual_seed(0)
    ref_model = FluxGNN(input_dim=4, hidden_dim=128, num_layers=4).eval()
    weights = {k: v.numpy().copy() for k, v in ref_model.state_dict().items()}
    out = {}
    for tag, nx, dt, radius, steps, every, n_ics in (("c2", 64, 1e-3, 3, 1000, 100, 4), ("c3", 1024, 3e-4, 2, 300, 100, 2)):
        base = BaselineSolver(nx=nx, dt=dt)
        g = P.Grid(nx=nx, dt=dt)
        ics = np.stack([P.stable_initial_condition(g, 50 + s) for s in range(n_ics)])
        s32, s64 = [], []
        for i, ic in enumerate(ics):
            r32 = reference_radius_run(ref_model, base, ic, steps, radius, every)
            assert np.isfinite(r32).all()
            if i == 0:          # the port walks the same trajectory bit for bit
                same(P.hybrid_run(weights, ic, g, steps, radius=radius)[::every], r32, f"hybrid_run {tag} r={radius} {steps} steps")
            s32.append(r32)
            s64.append(P.hybrid_run(weights, ic, g, steps, radius=radius, dtype=torch.float64)[::every])
        s32, s64 = np.stack(s32), np.stack(s64)
        print(f"      {tag}: reference fp32 vs fp64 after {steps} steps (n,u,E):", P.rel_err(s32[:, -1], s64[:, -1]))
        out.update({f"{tag}_ics": ics, f"{tag}_ref_fp32": s32, f"{tag}_fp64": s64, f"{tag}_dt": dt, f"{tag}_nx": nx,
                    f"{tag}_radius": radius, f"{tag}_steps": steps, f"{tag}_every": every})
    np.savez_compressed(os.path.join(OUT, "g6b_long_rollout_radius.npz"), **out)
    print("  wrote g6b_long_rollout_radius.npz")


def generic_architecture_goldens():
    """G10: the reference FluxGNN at architectures other than MODEL_CONFIG's -- (4, 64, 3) of examples/smoke_test.py:50-56,
    the class default (2, 32, 2) (src/flux_gnn.py:11), (4, 16, 1) -- forward on ring graphs, and a 5-step hybrid
    rollout with the (4, 64, 3) model driven with reference objects.  `--only-g10` regenerates just this file."""
    torch.set_num_threads(1)
    out = {}
    for tag, (F, H, L), seed in (("f4h64l3", (4, 64, 3), 5), ("f2h32l2", (2, 32, 2), 6), ("f4h16l1", (4, 16, 1), 7)):
        torch.manual_seed(seed)
        ref_model = FluxGNN(input_dim=F, hidden_dim=H, num_layers=L).eval()
        w = {k: v.numpy().copy() for k, v in ref_model.state_dict().items()}
        mine = P.init_weights(seed, F, H, L)
        for key in w:
            same(mine[key], w[key], f"init_weights[{tag}][{key}]")
        for nx in (64, 300):
            base = BaselineSolver(nx=nx)
            state = base.initial_condition(seed=31)
            nf4, _ = build_chain_graph(state, base.x)
            nf = nf4[:, :F].contiguous()                      # F = 2: features [n, u]
            out[f"{tag}_feats_nx{nx}"] = nf.numpy()
            for r in (1, 2, 3):
                edges = torch.from_numpy(P.ring_edges(nx, r))
                with torch.no_grad():
                    flux = ref_model(nf, edges).numpy()
                same(P.fluxgnn_forward(w, nf.numpy(), edges.numpy()), flux, f"forward {tag} nx={nx} r={r}")
                out[f"{tag}_flux_nx{nx}_r{r}"] = flux
        if F == 4:
            base = BaselineSolver(nx=64, dt=5e-3)
            ics = np.stack([base.initial_condition(seed=s) for s in (0, 1)])
            for r in (1, 2):
                runs = np.stack([reference_radius_run(ref_model, base, ic, 5, r, 1) for ic in ics])      # [2,6,3,64]
                same(P.hybrid_run(w, ics[0], P.Grid(nx=64, dt=5e-3), 5, radius=r), runs[0], f"hybrid_run {tag} r={r}")
                out[f"{tag}_rollout5_r{r}"] = runs
            out[f"{tag}_ics"] = ics
    np.savez_compressed(os.path.join(OUT, "g10_generic_architectures.npz"), **out)
    print("  wrote g10_generic_architectures.npz")


def comparison_goldens():
    """G9: the reference's PureGNN and PINN classes, seeded, against the port; their
    outputs frozen.  `python -m oracle.make_golden --only-g9` regenerates just this file."""
    sys.path.insert(0, os.path.join(REF, "scripts", "training"))
    from train_pinn import PINN                               # noqa: E402  (reference)
    from train_pure_gnn import PureGNN                        # noqa: E402  (reference)
    torch.set_num_threads(1)
    g9 = {}
    base = BaselineSolver(nx=64)
    ics = np.stack([base.initial_condition(seed=s) for s in range(4)])
    x32 = base.x.astype(np.float32)
    for hidden, layers in ((64, 3), (128, 4)):                # class default / the timing benchmark's size
        torch.manual_seed(7)
        ref_model = PureGNN(input_dim=4, hidden_dim=hidden, num_layers=layers).eval()
        w = {k: v.numpy().copy() for k, v in ref_model.state_dict().items()}
        mine = P.init_pure_gnn_weights(7, 4, hidden, layers)
        for key in w:
            same(mine[key], w[key], f"init_pure_gnn_weights[{hidden}][{key}]")
        tag = f"pgnn{hidden}"
        with torch.no_grad():
            feats, ei = build_chain_graph(ics[0], base.x, "cpu")
            delta = ref_model(feats, ei).numpy()
            same(P.pure_gnn_forward(w, feats.numpy(), ei.numpy()), delta, f"{tag} forward")
            g9[f"{tag}_delta"] = delta
            finals = []
            for ic in ics:                                    # the loop of benchmark_timing.py:129-143, 10 steps
                state = ic.copy()
                for _ in range(10):
                    _, ei = build_chain_graph(state, base.x, "cpu")
                    nf = torch.cat([torch.FloatTensor(state).permute(1, 0), torch.FloatTensor(x32).unsqueeze(1)], dim=1)
                    state = (torch.FloatTensor(state).permute(1, 0) + ref_model(nf, ei)).permute(1, 0).numpy()
                finals.append(state)
            same(P.pure_gnn_rollout(w, ics[0], x32, 10), finals[0], f"{tag} 10-step rollout")
            g9[f"{tag}_rollout10"] = np.stack(finals)
    torch.manual_seed(11)
    ref_pinn = PINN(input_dim=3 * 64, hidden_dim=256, num_layers=4).eval()
    wp = {k: v.numpy().copy() for k, v in ref_pinn.state_dict().items()}
    mine = P.init_pinn_weights(11, 192, 256, 4)
    for key in wp:
        same(mine[key], wp[key], f"init_pinn_weights[{key}]")
    with torch.no_grad():
        out = ref_pinn(torch.from_numpy(ics)).numpy()
        same(P.pinn_forward(wp, ics), out, "pinn forward")
        g9["pinn_step"] = out
        state = torch.from_numpy(ics[:1])
        for _ in range(10):                                   # benchmark_timing.py:186-189
            state = ref_pinn(state)
        g9["pinn_rollout10"] = state.numpy()
    g9["ics"] = ics
    np.savez_compressed(os.path.join(OUT, "g9_comparison_models.npz"), **g9)
    print("  wrote g9_comparison_models.npz")


def training_rollout_goldens():
    """G11: the reference's multi-step TRAINING rollout under its own autograd --
    scripts/training/train_ablation.py:172-206 driven with the reference's FluxGNN, build_chain_graph and
    that script's solve_poisson_np (detached, through numpy, as there) -- for the fused differentiable step
    (HybridSolver.step_with_grad).  The loss is the script's multi-step energy term (:204-205) plus seeded
    linear functionals of every step's face flux and of the final n, u, so that every gradient path carries
    signal.  `python -m oracle.make_golden --only-g11` regenerates just this file."""
    sys.path.insert(0, os.path.join(REF, "scripts", "training"))
    from train_ablation import solve_poisson_np               # noqa: E402  (reference)
    torch.set_num_threads(1)
    w = dict(np.load(os.path.join(OUT, "weights_seed0.npz")))
    g11 = {}
    for tag, nx, radius, dt, steps in (("nx64_r1", 64, 1, 5e-3, 3), ("nx64_r3", 64, 3, 5e-3, 3), ("nx256_r2", 256, 2, 1e-3, 2)):
        base = BaselineSolver(nx=nx, dt=dt)
        dx = base.dx
        ref_model = FluxGNN(input_dim=4, hidden_dim=128, num_layers=4)
        ref_model.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
        ref_model.train()
        edges = torch.from_numpy(P.ring_edges(nx, radius))
        ics = np.stack([base.initial_condition(seed=40 + s) for s in range(2)]).astype(np.float32)
        rs = np.random.RandomState(7 + nx + radius)
        cot_face = rs.randn(2, steps, nx).astype(np.float32)
        cot_n = rs.randn(2, nx).astype(np.float32)
        cot_u = rs.randn(2, nx).astype(np.float32)
        x = torch.from_numpy(base.x.astype(np.float32))
        loss = 0.0
        st0 = [torch.from_numpy(ic).requires_grad_(True) for ic in ics]
        finals, faces = [], []
        for b in range(2):
            state_roll = st0[b]
            energies = []
            for k in range(steps):                            # train_ablation.py:176-201
                n_r, u_r, E_r = state_roll[0], state_roll[1], state_roll[2]
                energies.append(0.5 * torch.mean(u_r ** 2))
                node_f_r, _ = build_chain_graph(state_roll, x, device="cpu")
                flux_edge_r = ref_model(node_f_r, edges)
                F_r = 0.5 * (flux_edge_r[:nx] + flux_edge_r[nx:2 * nx])
                n_next_r = n_r - (dt / dx) * (F_r - torch.roll(F_r, 1))
                F_u_r = 0.5 * u_r * u_r
                u_next_r = (u_r - (dt / dx) * (F_u_r - torch.roll(F_u_r, 1))) + dt * E_r
                E_next_r = torch.from_numpy(solve_poisson_np(n_next_r.detach().cpu().numpy(), 1.0, dx))
                state_roll = torch.stack([n_next_r, u_next_r, E_next_r], dim=0)
                loss = loss + (F_r * torch.from_numpy(cot_face[b, k])).sum()
                faces.append(F_r.detach().numpy())
            energies = torch.stack(energies)
            loss = loss + torch.mean((energies - energies[0]) ** 2)                 # :204-205
            loss = loss + (state_roll[0] * torch.from_numpy(cot_n[b])).sum() + (state_roll[1] * torch.from_numpy(cot_u[b])).sum()
            finals.append(state_roll.detach().numpy())
        loss.backward()
        g11[f"{tag}_ics"] = ics
        g11[f"{tag}_cot_face"], g11[f"{tag}_cot_n"], g11[f"{tag}_cot_u"] = cot_face, cot_n, cot_u
        g11[f"{tag}_final"] = np.stack(finals)
        g11[f"{tag}_faces"] = np.stack(faces).reshape(2, steps, nx)
        g11[f"{tag}_loss"] = np.float64(loss.item())
        g11[f"{tag}_dstate0"] = np.stack([s.grad.numpy() for s in st0])
        for name, p_ in ref_model.named_parameters():
            g = p_.grad.numpy()
            if g.size <= 1024:
                g11[f"{tag}_grad_{name}"] = g
            else:
                g11[f"{tag}_gradnorm_{name}"] = np.float64(np.linalg.norm(g.astype(np.float64)))
                g11[f"{tag}_gradcorner_{name}"] = g[:8, :8].copy()
                g11[f"{tag}_gradrows_{name}"] = g[[5, 77], :].copy()
        print(f"  ok  reference training rollout {tag}: loss {loss.item():.6e}")
    np.savez_compressed(os.path.join(OUT, "g11_training_rollout.npz"), **g11)
    print("  wrote g11_training_rollout.npz")


if __name__ == "__main__":
    if "--only-g9" in sys.argv:
        comparison_goldens()
    elif "--only-g11" in sys.argv:
        training_rollout_goldens()
    elif "--only-g6b" in sys.argv:
        long_rollout_goldens()
    elif "--only-g10" in sys.argv:
        generic_architecture_goldens()
    else:
        main()
