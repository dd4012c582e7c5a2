"""Batched closed-form CPU restatement of the hybrid step (oracle; test infrastructure).

On the radius-r ring the reference's gather / index_add_ / bincount message
passing (src/flux_gnn.py:53-60) is, for every node i,

    h'_i = relu(W[:, :H] h_i + W[:, H:] * (1/2r) sum_{k=1..r}(h_{i-k} + h_{i+k}) + b)

(indices periodic; every node has exactly 2r incoming edges, repeated when
nx <= 2r, and bincount counts the repeats too), and the edge readout
(src/flux_gnn.py:63-66) on edge (row a, col b) is  w2 . relu(W1[:, :H] h_a +
W1[:, H:] h_b + b1) + b2.  This file evaluates exactly that with dense torch
ops over a whole batch [B,3,nx]; tests check it against oracle/ref_port.py
(and so against the reference) before using it on shapes where the unbatched
port would take minutes.  It is also the strongest CPU baseline bench.py can
time (all host threads, one BLAS call per layer).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
import this module.
"""
from __future__ import annotations

import numpy as np
import torch

from .ref_port import N0, num_layers


def _neighbour_mean(h: torch.Tensor, radius: int) -> torch.Tensor:
    """(1/2r) sum_{k=1..r} (h_{i-k} + h_{i+k}) along dim -2 of h[B,nx,H]."""
    acc = torch.zeros_like(h)
    for hop in range(1, radius + 1):
        acc = acc + torch.roll(h, hop, dims=-2) + torch.roll(h, -hop, dims=-2)
    return acc / float(2 * radius)


def node_embeddings(w: dict, state: torch.Tensor, x: torch.Tensor, radius: int) -> torch.Tensor:
    """h^L[B,nx,H] after the input MLP and all message-passing layers."""
    dtype = state.dtype
    lin = torch.nn.functional.linear
    tw = lambda key: torch.as_tensor(w[key]).to(dtype)
    feats = torch.stack([state[:, 0], state[:, 1], state[:, 2], x.to(dtype).expand_as(state[:, 0])], dim=-1)
    h = torch.relu(lin(feats, tw("input_mlp.0.weight"), tw("input_mlp.0.bias")))
    for layer in range(num_layers(w)):
        cat = torch.cat([h, _neighbour_mean(h, radius)], dim=-1)
        h = torch.relu(lin(cat, tw(f"update_mlps.{layer}.0.weight"), tw(f"update_mlps.{layer}.0.bias")))
    return h


def edge_fluxes(w: dict, state: torch.Tensor, x: torch.Tensor, radius: int, hops: int | None = None) -> torch.Tensor:
    """Directed-edge fluxes [B, 2*hops*nx] in ring_edges() order
    (blocks [i->i+k], [i+k->i] for k = 1..hops)."""
    dtype = state.dtype
    hops = radius if hops is None else hops
    h = node_embeddings(w, state, x, radius)
    hid = h.shape[-1]
    w1 = torch.as_tensor(w["edge_mlp.0.weight"]).to(dtype)
    b1 = torch.as_tensor(w["edge_mlp.0.bias"]).to(dtype)
    w2 = torch.as_tensor(w["edge_mlp.2.weight"]).to(dtype)[0]
    b2 = torch.as_tensor(w["edge_mlp.2.bias"]).to(dtype)[0]
    p = h @ w1[:, :hid].T          # row-endpoint half
    q = h @ w1[:, hid:].T          # col-endpoint half
    blocks = []
    for hop in range(1, hops + 1):
        fwd = torch.relu(p + torch.roll(q, -hop, dims=-2) + b1) @ w2 + b2     # row i, col i+k
        bwd = torch.relu(torch.roll(p, -hop, dims=-2) + q + b1) @ w2 + b2     # row i+k, col i
        blocks += [fwd, bwd]
    return torch.cat(blocks, dim=-1)


def poisson(n: torch.Tensor, k: torch.Tensor) -> torch.Tensor:
    """Batched src/baseline_solver.py:59-68: forward FFT in the input precision,
    spectral multiply and inverse in complex128, real part, cast back."""
    rho = n - N0
    spec = torch.fft.fft(rho).to(torch.complex128)
    kk = k.to(torch.float64)
    mult = torch.zeros_like(kk)
    nz = kk != 0
    mult[nz] = 1.0 / kk[nz]
    return torch.fft.ifft(1j * spec * mult).real.to(n.dtype)


def hybrid_step(w: dict, state: torch.Tensor, x, k, dt: float, dx: float, radius: int = 1) -> torch.Tensor:
    """Batched src/hybrid_solver.py:34-64 on state[B,3,nx] (fp32 or fp64)."""
    with torch.no_grad():
        x = torch.as_tensor(np.asarray(x, dtype=np.float32))
        k = torch.as_tensor(np.asarray(k))
        nx = state.shape[-1]
        fl = edge_fluxes(w, state, x, radius, hops=1)
        face = 0.5 * (fl[:, :nx] + fl[:, nx:])
        c = float(np.float32(dt / dx))
        dt32 = float(np.float32(dt))
        n, u, E = state[:, 0], state[:, 1], state[:, 2]
        n_new = n - c * (face - torch.roll(face, 1, dims=-1))
        fu = 0.5 * u * u
        u_new = (u - c * (fu - torch.roll(fu, 1, dims=-1))) + dt32 * E
        return torch.stack([n_new, u_new, poisson(n_new, k)], dim=1)


def training_rollout(w: dict, state: torch.Tensor, x, k, dt: float, dx: float, n_steps: int, radius: int = 1):
    """The reference's multi-step TRAINING rollout, scripts/training/train_ablation.py:172-206, batched and
    differentiable: per step  model -> face flux -> n', u' with autograd edges -> field solve of n' DETACHED (it goes
    through numpy there, :198-200).  Returns (list of the n_steps + 1 states, list of the n_steps face fluxes); `w` may
    hold tensors that require grad, `state` too.  The checker of HybridSolver.rollout_with_grad."""
    x = torch.as_tensor(np.asarray(x, dtype=np.float32))
    k = torch.as_tensor(np.asarray(k))
    nx = state.shape[-1]
    c, dt32 = float(np.float32(dt / dx)), float(np.float32(dt))
    states, faces = [state], []
    for _ in range(n_steps):
        cur = states[-1]
        fl = edge_fluxes(w, cur, x.to(cur.dtype) if cur.dtype == torch.float64 else x, radius, hops=1)
        face = 0.5 * (fl[:, :nx] + fl[:, nx:])
        n, u, E = cur[:, 0], cur[:, 1], cur[:, 2]
        n_new = n - c * (face - torch.roll(face, 1, dims=-1))
        fu = 0.5 * u * u
        u_new = (u - c * (fu - torch.roll(fu, 1, dims=-1))) + dt32 * E
        with torch.no_grad():
            E_new = poisson(n_new.detach(), k)
        states.append(torch.stack([n_new, u_new, E_new], dim=1))
        faces.append(face)
    return states, faces


def hybrid_run(w, state0: torch.Tensor, x, k, dt, dx, n_steps: int, radius: int = 1, record_every: int = 0):
    """Rollout; returns the final state, or [1+T/record_every, B, 3, nx] when recording."""
    state = state0
    traj = [state0] if record_every else None
    for t in range(1, n_steps + 1):
        state = hybrid_step(w, state, x, k, dt, dx, radius)
        if record_every and t % record_every == 0:
            traj.append(state)
    return torch.stack(traj, dim=0) if record_every else state


def baseline_step(state: torch.Tensor, k, dt: float, dx: float, nu: float) -> torch.Tensor:
    """Batched src/baseline_solver.py:80-101."""
    with torch.no_grad():
        k = torch.as_tensor(np.asarray(k))
        c = float(np.float32(dt / dx))
        dt32, nu32, dx2 = float(np.float32(dt)), float(np.float32(nu)), float(np.float32(dx ** 2))
        n, u, E = state[:, 0], state[:, 1], state[:, 2]
        fn = n * u
        n_new = n - c * (fn - torch.roll(fn, 1, dims=-1))
        fu = 0.5 * u * u
        u_adv = u - c * (fu - torch.roll(fu, 1, dims=-1))
        lap = (torch.roll(u, -1, dims=-1) - 2 * u + torch.roll(u, 1, dims=-1)) / dx2
        u_new = u_adv + dt32 * (E + nu32 * lap)
        return torch.stack([n_new, u_new, poisson(n_new, k)], dim=1)
