"""Unbatched CPU restatement of the reference hot path (oracle; test infrastructure).

Each function names the reference lines (relative to /root/reference) whose
arithmetic it restates.  Weights travel as a plain ``dict`` of numpy arrays
keyed exactly like ``FluxGNN.state_dict()`` (src/flux_gnn.py:17-38):

    input_mlp.0.{weight[H,F],bias[H]}
    update_mlps.{l}.0.{weight[H,2H],bias[H]}   l = 0..L-1
    edge_mlp.0.{weight[H,2H],bias[H]}
    edge_mlp.2.{weight[1,H],bias[1]}

The fp32 mode uses the same torch/numpy primitives in the same order as the
reference, so on one machine it is bit-identical to it (checked by
oracle/make_golden.py).  ``dtype=torch.float64`` gives the fp64 restatement
used to measure the reference's own fp32 noise floor.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
import this module.
"""
from __future__ import annotations

import math

import numpy as np
import torch

N0 = 1.0  # background density, src/baseline_solver.py:20


# --------------------------------------------------------------------------
# grid, graph
# --------------------------------------------------------------------------
class Grid:
    """Cell-centred periodic grid + wavenumbers (src/baseline_solver.py:14-27)."""

    def __init__(self, nx=64, length=2 * math.pi, dt=5e-3, nu=1e-3):
        self.nx = int(nx)
        self.length = float(length)
        self.dx = self.length / self.nx
        self.dt = float(dt)
        self.nu = float(nu)
        self.x = np.linspace(0.5 * self.dx, self.length - 0.5 * self.dx, self.nx)
        self.k = 2.0 * np.pi * np.fft.fftfreq(self.nx, d=self.dx)


def ring_edges(nx: int, radius: int = 1) -> np.ndarray:
    """Directed edges of the periodic chain.

    radius 1 reproduces src/graph_constructor.py:34-38 column for column:
    columns [0,nx) are (row i, col i+1), columns [nx,2nx) are (row i+1, col i).
    For radius r > 1 (an extension: the reference never builds it)
    the same pair of blocks is appended for every hop distance k = 2..r, so the
    first 2*nx columns never change.
    """
    src = np.arange(nx, dtype=np.int64)
    rows, cols = [], []
    for hop in range(1, radius + 1):
        dst = (src + hop) % nx
        rows += [src, dst]
        cols += [dst, src]
    return np.stack([np.concatenate(rows), np.concatenate(cols)], axis=0)


def node_features(state: np.ndarray, x: np.ndarray) -> np.ndarray:
    """AoS node features [n, u, E, x] as float32 (src/graph_constructor.py:14-32)."""
    st = np.asarray(state)
    cols = [st[c].astype(np.float32) for c in range(3)]
    cols.append(np.asarray(x).astype(np.float32))
    return np.stack(cols, axis=-1)


# --------------------------------------------------------------------------
# FluxGNN forward on an arbitrary edge list
# --------------------------------------------------------------------------
def num_layers(w: dict) -> int:
    return sum(1 for key in w if key.startswith("update_mlps.") and key.endswith(".weight"))


def _t(a, dtype):
    return torch.as_tensor(np.asarray(a)).to(dtype)


def fluxgnn_forward(w: dict, feats, edge_index, dtype=torch.float32) -> np.ndarray:
    """Generic message-passing forward (src/flux_gnn.py:40-67).

    h0 = relu(X W_in^T + b)                                   (:49)
    per layer: agg[row] += h[col]; agg /= max(deg,1);         (:55-59)
               h = relu([h, agg] W_l^T + b_l)                 (:60)
    flux_e = w2 . relu([h_row, h_col] W_e1^T + b_e1) + b2     (:63-66)
    """
    with torch.no_grad():
        lin = torch.nn.functional.linear
        h = torch.relu(lin(_t(feats, dtype), _t(w["input_mlp.0.weight"], dtype),
                           _t(w["input_mlp.0.bias"], dtype)))
        ei = torch.as_tensor(np.asarray(edge_index), dtype=torch.long)
        row, col = ei[0], ei[1]
        n_nodes = h.shape[0]
        for layer in range(num_layers(w)):
            acc = torch.zeros_like(h)
            acc.index_add_(0, row, h[col])
            deg = torch.bincount(row, minlength=n_nodes).clamp_min(1).to(dtype).unsqueeze(-1)
            acc = acc / deg
            h = torch.relu(lin(torch.cat([h, acc], dim=-1),
                               _t(w[f"update_mlps.{layer}.0.weight"], dtype),
                               _t(w[f"update_mlps.{layer}.0.bias"], dtype)))
        pair = torch.cat([h[row], h[col]], dim=-1)
        hid = torch.relu(lin(pair, _t(w["edge_mlp.0.weight"], dtype), _t(w["edge_mlp.0.bias"], dtype)))
        out = lin(hid, _t(w["edge_mlp.2.weight"], dtype), _t(w["edge_mlp.2.bias"], dtype)).squeeze(-1)
    return out.numpy()


# --------------------------------------------------------------------------
# field solve
# --------------------------------------------------------------------------
def solve_poisson(n: np.ndarray, k: np.ndarray, pinned_numpy: bool = False) -> np.ndarray:
    """Spectral field solve E = Re ifft(i * fft(n - n0) / k), zero mean mode
    (src/baseline_solver.py:59-68).  The Nyquist bin is annihilated by Re().

    ``pinned_numpy`` upcasts rho to float64 before the forward FFT, which is
    what numpy 1.26.4 (the reference's pin, requirements.txt:5) does; numpy >= 2
    keeps a float32 input in complex64 for the forward transform.
    """
    rho = n - N0
    if pinned_numpy:
        rho = rho.astype(np.float64)
    spec = np.fft.fft(rho)
    out = np.zeros_like(spec, dtype=complex)
    nz = k != 0
    out[nz] = 1j * spec[nz] / k[nz]
    return np.real(np.fft.ifft(out)).astype(np.float32)


def poisson_kernel(nx: int, length: float) -> np.ndarray:
    """The same operator as a circular-convolution kernel g (float64):
    solve_poisson(n) == circular_conv(g, n - n0).  g = Re ifft(i/k), g_hat(0)=0."""
    dx = length / nx
    k = 2.0 * np.pi * np.fft.fftfreq(nx, d=dx)
    ghat = np.zeros(nx, dtype=complex)
    nz = k != 0
    ghat[nz] = 1j / k[nz]
    return np.real(np.fft.ifft(ghat))


# --------------------------------------------------------------------------
# solvers
# --------------------------------------------------------------------------
def hybrid_step(w: dict, state: np.ndarray, g: Grid, radius: int = 1,
                dtype=torch.float32) -> np.ndarray:
    """One hybrid step (src/hybrid_solver.py:34-64): GNN interface flux for the
    continuity equation, conservative left-differenced u^2/2 plus dt*E for the
    momentum equation (no viscosity), then the field solve."""
    if dtype == torch.float64:
        return _hybrid_step_f64(w, state, g, radius)
    n, u, E = state
    flux = fluxgnn_forward(w, node_features(state, g.x), ring_edges(g.nx, radius))
    nx = g.nx
    face = 0.5 * (flux[:nx] + flux[nx:2 * nx]).astype(np.float32)     # :45-48
    c = g.dt / g.dx
    n_new = n - c * (face - np.roll(face, 1))                          # :51-52
    fu = 0.5 * u * u                                                   # :55
    u_new = (u - c * (fu - np.roll(fu, 1))) + g.dt * E                 # :56-58
    E_new = solve_poisson(n_new, g.k)                                  # :61
    return np.stack([n_new, u_new, E_new], axis=0).astype(np.float32)  # :63


def _hybrid_step_f64(w, state, g, radius):
    st = np.asarray(state, dtype=np.float64)
    n, u, E = st
    feats = np.stack([n, u, E, g.x.astype(np.float32).astype(np.float64)], axis=-1)
    flux = fluxgnn_forward(w, feats, ring_edges(g.nx, radius), dtype=torch.float64)
    nx = g.nx
    face = 0.5 * (flux[:nx] + flux[nx:2 * nx])
    c = float(np.float32(g.dt / g.dx))
    dt = float(np.float32(g.dt))
    n_new = n - c * (face - np.roll(face, 1))
    fu = 0.5 * u * u
    u_new = (u - c * (fu - np.roll(fu, 1))) + dt * E
    spec = np.fft.fft(n_new - N0)
    out = np.zeros_like(spec)
    nz = g.k != 0
    out[nz] = 1j * spec[nz] / g.k[nz]
    return np.stack([n_new, u_new, np.real(np.fft.ifft(out))], axis=0)


def hybrid_run(w, state0, g, n_steps=40, radius=1, dtype=torch.float32):
    """T-step loop returning [T+1,3,nx] including state0 (src/hybrid_solver.py:66-73)."""
    cast = np.float32 if dtype == torch.float32 else np.float64
    state = np.asarray(state0).astype(cast)
    out = [state]
    for _ in range(n_steps):
        state = hybrid_step(w, state, g, radius, dtype)
        out.append(state)
    return np.stack(out, axis=0)


def baseline_step(state: np.ndarray, g: Grid, return_flux: bool = False):
    """Classical upwind / forward-Euler step with viscosity
    (src/baseline_solver.py:70-101)."""
    n, u, E = state
    c = g.dt / g.dx
    fn = (n * u).astype(np.float32)                                    # :70-71, :84
    n_new = n - c * (fn - np.roll(fn, 1))                              # :85-86
    fu = (0.5 * u * u).astype(np.float32)                              # :73-74, :89
    u_adv = u - c * (fu - np.roll(fu, 1))                              # :90-91
    lap = (np.roll(u, -1) - 2 * u + np.roll(u, 1)) / (g.dx ** 2)       # :76-78
    u_new = u_adv + g.dt * (E + g.nu * lap)                            # :94
    E_new = solve_poisson(n_new, g.k)                                  # :96
    new = np.stack([n_new, u_new, E_new], axis=0).astype(np.float32)
    return (new, fn) if return_flux else new


def baseline_run(state0, g: Grid, n_steps=10, record_flux=True):
    """src/baseline_solver.py:103-118."""
    state = np.asarray(state0).astype(np.float32)
    states, fluxes = [state], []
    for _ in range(n_steps):
        state, fn = baseline_step(state, g, return_flux=True)
        states.append(state)
        fluxes.append(fn)
    return np.stack(states, axis=0), (np.stack(fluxes, axis=0) if record_flux else None)


# --------------------------------------------------------------------------
# inputs
# --------------------------------------------------------------------------
def initial_condition(g: Grid, seed=None) -> np.ndarray:
    """The reference's random-mode initial condition; the draw order is part of
    the contract (src/baseline_solver.py:29-57)."""
    rng = np.random.RandomState(seed)
    n = np.full(g.nx, N0, dtype=np.float32)
    for _ in range(rng.randint(3, 6)):
        mode = rng.randint(1, 6)
        amp = 0.15 + 0.15 * rng.rand()
        phase = 2 * np.pi * rng.rand()
        n += amp * np.sin(mode * g.x + phase).astype(np.float32)
    u = np.zeros(g.nx, dtype=np.float32)
    for _ in range(2):
        mode = rng.randint(1, 6)
        amp = 0.1 + 0.1 * rng.rand()
        phase = 2 * np.pi * rng.rand()
        u += amp * np.cos(mode * g.x + phase).astype(np.float32)
    u += 0.05 * rng.randn(g.nx).astype(np.float32)
    E = solve_poisson(n, g.k)
    return np.stack([n, u, E], axis=0).astype(np.float32)


def stable_initial_condition(g: Grid, seed: int) -> np.ndarray:
    """Smooth low-amplitude recipe for long rollouts (NOT in the
    reference, whose own ICs go non-finite after ~200 steps):
    4 density sine modes and 2 velocity cosine modes, amplitude in [0, 0.1),
    mode number in 1..5, uniform phase, no white noise."""
    rng = np.random.RandomState(seed)
    n = np.full(g.nx, N0, dtype=np.float64)
    for _ in range(4):
        mode, amp, phase = rng.randint(1, 6), 0.1 * rng.rand(), 2 * np.pi * rng.rand()
        n += amp * np.sin(mode * g.x + phase)
    u = np.zeros(g.nx, dtype=np.float64)
    for _ in range(2):
        mode, amp, phase = rng.randint(1, 6), 0.1 * rng.rand(), 2 * np.pi * rng.rand()
        u += amp * np.cos(mode * g.x + phase)
    n = n.astype(np.float32)
    return np.stack([n, u.astype(np.float32), solve_poisson(n, g.k)], axis=0)


def init_weights(seed: int = 0, input_dim: int = 4, hidden: int = 128, layers: int = 4) -> dict:
    """Random-init weights with torch's default nn.Linear initialisation, created
    in the reference's construction order (src/flux_gnn.py:17-38) so that
    ``torch.manual_seed(seed); FluxGNN(input_dim, hidden, layers)`` gives the
    same tensors."""
    torch.manual_seed(seed)
    mods = [("input_mlp.0", torch.nn.Linear(input_dim, hidden))]
    mods += [(f"update_mlps.{l}.0", torch.nn.Linear(2 * hidden, hidden)) for l in range(layers)]
    mods += [("edge_mlp.0", torch.nn.Linear(2 * hidden, hidden)), ("edge_mlp.2", torch.nn.Linear(hidden, 1))]
    w = {}
    for name, m in mods:
        w[name + ".weight"] = m.weight.detach().numpy().copy()
        w[name + ".bias"] = m.bias.detach().numpy().copy()
    return w


# --------------------------------------------------------------------------
# diagnostics used by the reference's evaluation scripts
# --------------------------------------------------------------------------
def rel_err(a, b) -> np.ndarray:
    """Per-channel max-norm relative error ||a-b||_inf / ||b||_inf over the last
    axis set (the parity metric of SURVEY 8d)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    a = np.moveaxis(a, -2, 0).reshape(a.shape[-2], -1)
    b = np.moveaxis(b, -2, 0).reshape(b.shape[-2], -1)
    return np.abs(a - b).max(axis=1) / np.maximum(np.abs(b).max(axis=1), 1e-30)


def energy(states) -> np.ndarray:
    """0.5*mean(u^2+E^2) per stored step (scripts/evaluation/evaluate_all.py:134-135)."""
    s = np.asarray(states, dtype=np.float64)
    return 0.5 * np.mean(s[..., 1, :] ** 2 + s[..., 2, :] ** 2, axis=-1)


def charge(states) -> np.ndarray:
    """mean(n) per stored step (scripts/evaluation/evaluate_all.py:140-141)."""
    return np.mean(np.asarray(states, dtype=np.float64)[..., 0, :], axis=-1)


def compute_metrics(states_pred, states_true) -> dict:
    """Evaluation metrics of scripts/evaluation/evaluate_all.py:118-159 for trajectories
    [T,3,nx]: per-step channel MSE, energy / charge and their drifts, final values."""
    sp, st = np.asarray(states_pred), np.asarray(states_true)
    mse = [np.mean((sp[:, c] - st[:, c]) ** 2, axis=1) for c in range(3)]
    total = mse[0] + mse[1] + mse[2]
    e_pred = 0.5 * np.mean(sp[:, 1] ** 2 + sp[:, 2] ** 2, axis=1)
    e_true = 0.5 * np.mean(st[:, 1] ** 2 + st[:, 2] ** 2, axis=1)
    c_pred, c_true = np.mean(sp[:, 0], axis=1), np.mean(st[:, 0], axis=1)
    return {"mse_n": mse[0], "mse_u": mse[1], "mse_E": mse[2], "mse_total": total,
            "energy_drift_pred": np.abs(e_pred - e_pred[0]), "energy_drift_true": np.abs(e_true - e_true[0]),
            "charge_drift_pred": np.abs(c_pred - c_pred[0]), "charge_drift_true": np.abs(c_true - c_true[0]),
            "final_mse": float(total[-1]), "mean_mse": float(np.mean(total)),
            "final_energy_drift": float(np.abs(e_pred - e_pred[0])[-1]),
            "final_charge_drift": float(np.abs(c_pred - c_pred[0])[-1])}


def generate_dataset(nx=64, num_initial_conditions=20, steps_per_ic=30, dt=5e-3, nu=1e-3):
    """scripts/training/generate_data.py:12-54 without the file output: (state_t, flux_t, state_next)."""
    g = Grid(nx=nx, dt=dt, nu=nu)
    st, fl, nxt = [], [], []
    for ic in range(num_initial_conditions):
        states, fluxes = baseline_run(initial_condition(g, seed=ic), g, n_steps=steps_per_ic)
        st.append(states[:-1]); fl.append(fluxes); nxt.append(states[1:])
    return np.concatenate(st), np.concatenate(fl), np.concatenate(nxt)


# --------------------------------------------------------------------------
# the reference's comparison models
# --------------------------------------------------------------------------
def init_pure_gnn_weights(seed: int = 0, input_dim: int = 4, hidden: int = 64, layers: int = 3) -> dict:
    """``torch.manual_seed(seed); PureGNN(input_dim, hidden, layers).state_dict()`` in the reference's
    construction order (scripts/training/train_pure_gnn.py:37-58)."""
    torch.manual_seed(seed)
    mods = [("input_mlp.0", torch.nn.Linear(input_dim, hidden))]
    mods += [(f"update_mlps.{l}.0", torch.nn.Linear(2 * hidden, hidden)) for l in range(layers)]
    mods += [("output_mlp.0", torch.nn.Linear(hidden, hidden)), ("output_mlp.2", torch.nn.Linear(hidden, 3))]
    return {f"{name}.{p}": getattr(m, p).detach().numpy().copy() for name, m in mods for p in ("weight", "bias")}


def pure_gnn_forward(w: dict, feats, edge_index, dtype=torch.float32) -> np.ndarray:
    """PureGNN.forward, scripts/training/train_pure_gnn.py:60-76: delta_state [N,3]."""
    lin = lambda name, v: torch.nn.functional.linear(v, _t(w[name + ".weight"], dtype), _t(w[name + ".bias"], dtype))
    h = torch.tanh(lin("input_mlp.0", _t(feats, dtype)))
    ei = torch.as_tensor(np.asarray(edge_index), dtype=torch.long)
    src, dst = ei[0], ei[1]
    layers = sum(1 for k in w if k.startswith("update_mlps.") and k.endswith(".weight"))
    for l in range(layers):
        upd = torch.tanh(lin(f"update_mlps.{l}.0", torch.cat([h[src], h[dst]], dim=-1)))
        agg = torch.zeros_like(h)
        agg.index_add_(0, dst, upd)
        h = h + agg
    return lin("output_mlp.2", torch.tanh(lin("output_mlp.0", h))).numpy()


def pure_gnn_rollout(w: dict, state0: np.ndarray, x: np.ndarray, n_steps: int, dtype=torch.float32) -> np.ndarray:
    """The PureGNN loop of scripts/evaluation/benchmark_timing.py:129-143 for one IC: final [3,nx]."""
    state = np.asarray(state0, dtype=np.float32)
    ei = ring_edges(state.shape[1], 1)
    for _ in range(n_steps):
        feats = np.concatenate([state.T, np.asarray(x, dtype=np.float32)[:, None]], axis=1)
        delta = pure_gnn_forward(w, feats, ei, dtype)
        state = (state.T.astype(delta.dtype) + delta).T.astype(np.float32) if dtype == torch.float32 else (state.T + delta).T
    return state


def init_pinn_weights(seed: int = 0, input_dim: int = 192, hidden: int = 256, layers: int = 4) -> dict:
    """``torch.manual_seed(seed); PINN(input_dim, hidden, layers).state_dict()`` (scripts/training/train_pinn.py:38-48)."""
    torch.manual_seed(seed)
    dims = [input_dim] + [hidden] * (layers - 1) + [input_dim]
    w = {}
    for i in range(layers):
        m = torch.nn.Linear(dims[i], dims[i + 1])
        w[f"net.{2 * i}.weight"] = m.weight.detach().numpy().copy()
        w[f"net.{2 * i}.bias"] = m.bias.detach().numpy().copy()
    return w


def pinn_forward(w: dict, state, dtype=torch.float32) -> np.ndarray:
    """PINN.forward, scripts/training/train_pinn.py:50-61: state + net(flatten(state))."""
    s = _t(state, dtype)
    flat = s.reshape(*s.shape[:-2], -1)
    n = sum(1 for k in w if k.endswith(".weight"))
    v = flat
    for i in range(n):
        v = torch.nn.functional.linear(v, _t(w[f"net.{2 * i}.weight"], dtype), _t(w[f"net.{2 * i}.bias"], dtype))
        if i < n - 1:
            v = torch.tanh(v)
    return (s + v.reshape(s.shape)).numpy()
