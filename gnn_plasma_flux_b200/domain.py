"""Domain decomposition of ONE large periodic grid across GPUs (BASELINE.json configs[3]).

The reference has no distributed code; this is new work behind the same step arithmetic
(src/hybrid_solver.py:34-64).  Rank r of G owns the contiguous slab of S = nx/G cells
[r*S, (r+1)*S).  Per time step:

  1. ring halo exchange of raw state: H = L*radius + 1 cells per side (the receptive field of
     one step), 12 bytes per cell, two small point-to-point messages per neighbour
     -- instead of exchanging 512-byte hidden states after every layer;
  2. the fused tile kernel on the slab + ghost cells (fluxgnn_hybrid_slab_step): n', u';
  3. the field solve is global: all-gather n' (4 B/cell), every rank runs the same FFT solve
     and keeps its slab of E'.  The solve is replicated on purpose: at these sizes it costs
     well under 2 % of the GNN work of the step, and an all-gather is one collective instead of
     the four all-to-all transposes of a distributed FFT.

One process per GPU; `TorchDistComm` speaks torch.distributed (NCCL on GPUs, gloo in the CPU
tests).  `step_emulated` runs G virtual ranks inside one process on one GPU -- the way to
exercise the decomposition without G devices (separate processes that wait on one another
must not share a GPU).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .baseline_solver import BaselineSolver
from .grid import PeriodicGrid


class TorchDistComm:
    """Ring neighbours + all-gather over a torch.distributed process group."""

    def __init__(self, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def exchange_halos(self, left_edge: torch.Tensor, right_edge: torch.Tensor):
        """Send my first/last H cells to the left/right neighbour; return (left ghosts, right ghosts)."""
        left, right = (self.rank - 1) % self.world, (self.rank + 1) % self.world
        left_ghost, right_ghost = torch.empty_like(right_edge), torch.empty_like(left_edge)
        if self.world == 1:
            left_ghost.copy_(right_edge)
            right_ghost.copy_(left_edge)
            return left_ghost, right_ghost
        # order matters when left == right (two ranks): the peer's first receive must meet my first send
        ops = [dist.P2POp(dist.isend, right_edge.contiguous(), right, self.group),
               dist.P2POp(dist.isend, left_edge.contiguous(), left, self.group),
               dist.P2POp(dist.irecv, left_ghost, left, self.group),
               dist.P2POp(dist.irecv, right_ghost, right, self.group)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        return left_ghost, right_ghost

    def all_gather(self, t: torch.Tensor) -> torch.Tensor:
        out = torch.empty((self.world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        dist.all_gather(list(out.unbind(0)), t.contiguous(), group=self.group)   # views of `out`: no extra copy
        return out


class DomainDecomposedHybridSolver:
    """The hybrid step on one slab of a grid of `nx` cells split over `world` ranks."""

    def __init__(self, model, nx, length=2 * np.pi, dt=5e-3, graph_radius=1, rank=0, world=1, device="cuda",
                 precision="fp32", slab_fn=None, field_fn=None):
        if nx % world:
            raise ValueError(f"nx={nx} is not divisible by the number of ranks {world}")
        self.model, self.nx, self.length, self.dt = model, int(nx), float(length), float(dt)
        self.rank, self.world, self.device = rank, world, torch.device(device)
        self.radius = int(graph_radius)
        self.owned = self.nx // world
        self.halo = model.num_layers * self.radius + 1
        if self.owned < self.halo:
            raise ValueError(f"a slab of {self.owned} cells is narrower than the halo of {self.halo}")
        if precision != "fp32" and precision not in _lib.TC_PRECISIONS:
            raise ValueError(f"precision must be 'fp32' or one of {sorted(_lib.TC_PRECISIONS)}, got {precision!r}")
        self.precision = precision
        self.grid = PeriodicGrid(self.nx, self.length)
        idx = (rank * self.owned - self.halo + np.arange(self.owned + 2 * self.halo)) % self.nx
        self.x_ext = torch.as_tensor(self.grid.x[idx], dtype=torch.float32).to(self.device)   # GLOBAL positions
        self._slab_fn = slab_fn or self._cuda_slab
        self._field_fn = field_fn
        self._baseline = None

    # ---- local pieces -------------------------------------------------------------------
    def _cuda_slab(self, ext: torch.Tensor) -> torch.Tensor:
        tensor_path = self.precision != "fp32"
        packed = self.model.packed_weights(_lib.weight_layout(self.precision))
        B = ext.shape[0]
        with torch.cuda.device(self.device):
            out = torch.empty(B, 3, self.owned, dtype=torch.float32, device=self.device)
            stream = torch.cuda.current_stream(self.device).cuda_stream
            dx = self.length / self.nx
            _lib.check(_lib.lib().fluxgnn_hybrid_slab_step(
                packed.data_ptr(), self.model.num_layers, _lib.TC_PRECISIONS[self.precision] if tensor_path else 0,
                ext.data_ptr(), self.x_ext.data_ptr(), out.data_ptr(), B, self.owned, self.halo, self.radius,
                float(np.float32(self.dt / dx)), float(np.float32(self.dt)), stream), "fluxgnn_hybrid_slab_step")
        return out

    def advance_slab(self, ext: torch.Tensor) -> torch.Tensor:
        """ext [B,3,owned+2*halo] (ghosts included) -> [B,3,owned] with n', u' (E' not yet)."""
        if ext.shape[1:] != (3, self.owned + 2 * self.halo):
            raise ValueError(f"ext must be [B,3,{self.owned + 2 * self.halo}], got {tuple(ext.shape)}")
        return self._slab_fn(ext.to(torch.float32).contiguous())

    def field(self, n_full: torch.Tensor) -> torch.Tensor:
        """Global field solve E[B,nx] from the gathered density (src/baseline_solver.py:59-68)."""
        if self._field_fn is not None:
            return self._field_fn(n_full)
        if self._baseline is None:
            self._baseline = BaselineSolver(nx=self.nx, length=self.length, dt=self.dt, device=self.device)
        return self._baseline.solve_poisson(n_full.contiguous())

    # ---- one step, one process per rank ----------------------------------------------------
    def step(self, local: torch.Tensor, comm) -> torch.Tensor:
        """local [B,3,owned] -> new local state; `comm` provides exchange_halos / all_gather."""
        H, S = self.halo, self.owned
        left_ghost, right_ghost = comm.exchange_halos(local[..., :H].contiguous(), local[..., S - H:].contiguous())
        out = self.advance_slab(torch.cat([left_ghost, local, right_ghost], dim=-1))
        gathered = comm.all_gather(out[:, 0, :].contiguous())                    # [world,B,S]
        n_full = gathered.permute(1, 0, 2).reshape(local.shape[0], self.nx)
        out[:, 2, :] = self.field(n_full)[:, self.rank * S:(self.rank + 1) * S]
        return out


def split_slabs(state: torch.Tensor, world: int):
    """[B,3,nx] -> list of `world` contiguous slabs [B,3,nx/world]."""
    return [s.contiguous() for s in state.chunk(world, dim=-1)]


def step_emulated(solvers, locals_):
    """One decomposed step of G virtual ranks inside this process (solvers[r] built with rank=r,
    world=G): the halo exchange and the all-gather become tensor copies.  For single-GPU tests."""
    G = len(solvers)
    H, S = solvers[0].halo, solvers[0].owned
    exts = [torch.cat([locals_[(r - 1) % G][..., S - H:], locals_[r], locals_[(r + 1) % G][..., :H]], dim=-1)
            for r in range(G)]
    outs = [solvers[r].advance_slab(exts[r]) for r in range(G)]
    n_full = torch.cat([o[:, 0, :] for o in outs], dim=-1)
    for r in range(G):
        outs[r][:, 2, :] = solvers[r].field(n_full)[:, r * S:(r + 1) * S]
    return outs
