"""Domain decomposition of ONE large periodic grid across GPUs (BASELINE.json configs[3] and the
baseline-only row).

The reference has no distributed code; this is new work behind the same step arithmetic
(src/hybrid_solver.py:34-64, src/baseline_solver.py:80-101).  Rank r of G owns the contiguous slab of
S = nx/G cells [r*S, (r+1)*S) and keeps it permanently in the EXTENDED layout [B][3][S + 2H] (H ghost
cells per side), in two ping-pong buffers.  Per time step:

  1. ring halo exchange of raw state: H = L*radius + 1 cells per side for the hybrid step (the receptive
     field of one step; 12 bytes per cell instead of 512-byte hidden states after every layer), H = 4 for
     the classical step (its stencil needs 1; 4 keeps 128-bit loads aligned).  The two edges travel as one
     small message per neighbour and land in the ghost zones of the current buffer;
  2. the slab kernel (fluxgnn_hybrid_slab_step_ld / fluxgnn_baseline_slab_step) writes n', u' straight into
     the interior of the other buffer -- no concatenation, no slicing on the step path;
  3. the field solve is distributed (`field_solve="alltoall"`, the default wherever slabs and rank count are
     powers of two): two ICs travel as one complex
     signal, the nx-point transform is decimated in frequency over the rank index, four all-to-alls of
     4 bytes per cell each (include/fluxgnn.h, "distributed field solve"); every rank transforms only its own
     1/G of the spectrum.  `field_solve="allgather"` keeps round 1's variant (all-gather of n', the whole FFT
     replicated on every rank: G times the bytes and the arithmetic) as a cross-check.

Built with `fabric=SymmetricMemoryFabric()` the same solvers run the step over PEER MEMORY instead (`step_peer`,
`advance`): extended states and the gather buffer of the prefix-sum solve live in symmetric memory, the slab / field /
message kernels store the edge cells and messages straight into the neighbours' buffers, two signal-pad barriers order
a step, and pairs of steps replay as one CUDA graph.

One process per GPU; `TorchDistComm` speaks torch.distributed (NCCL on GPUs, gloo in the CPU tests).
`step_emulated` runs G virtual ranks inside one process on one GPU -- the way to exercise the
decomposition without G devices (separate processes that wait on one another must not share a GPU).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .baseline_solver import BaselineSolver
from .grid import PeriodicGrid


class TorchDistComm:
    """Ring neighbours, all-to-all and all-gather over a torch.distributed process group."""

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
            req.wait()               # NCCL: orders the current stream after the transfer, the host does not block
        return left_ghost, right_ghost

    def all_to_all(self, recv: torch.Tensor, send: torch.Tensor):
        """Flat buffers cut into `world` equal chunks: chunk q of `send` goes to rank q."""
        if self.world == 1:
            recv.copy_(send)
        else:
            dist.all_to_all_single(recv, send, group=self.group)

    def all_gather(self, t: torch.Tensor) -> torch.Tensor:
        out = torch.empty((self.world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        dist.all_gather(list(out.unbind(0)), t.contiguous(), group=self.group)   # views of `out`: no extra copy
        return out


# ------------------------------------------------------------------------------------------- peer memory
class _PeerBlock:
    """One symmetric allocation: the same number of bytes on every rank, every rank's copy addressable from here."""

    def __init__(self, fabric, local, views, bases_dev, barrier):
        self.fabric, self.local, self._views, self.bases_dev, self.barrier = fabric, local, views, bases_dev, barrier

    def view(self, rank: int, offset_bytes: int, shape, dtype=torch.float32) -> torch.Tensor:
        """Tensor over `shape` elements of rank `rank`'s copy, starting `offset_bytes` into it."""
        return self._views(rank, offset_bytes, tuple(shape), dtype)


class SymmetricMemoryFabric:
    """Peer memory of the ranks of a process group, one process per GPU: torch.distributed._symmetric_memory
    allocations (CUDA VMM, mapped into every rank over NVLink / NVSwitch) with their signal pads.  Halo cells and
    field-solve messages are then plain stores into the neighbours' memory, issued by the kernels that produce them
    (fluxgnn_hybrid_slab_step_peer, fluxgnn_baseline_slab_step_peer, fluxgnn_scan_slab_sums_peer,
    fluxgnn_scan_slab_field_peer; fluxgnn_peer_halo_push / fluxgnn_peer_allgather are the stand-alone forms) and ordered
    by the signal-pad barrier: no NCCL call on the step path."""

    def __init__(self, group=None, device=None, barrier_timeout_ms: int = 60000):
        """barrier_timeout_ms: a barrier that waits longer than this for a peer traps (a CUDA error on this rank) instead
        of spinning forever -- a rank that died must not hang the others' GPUs."""
        import torch.distributed._symmetric_memory as symm
        self._symm = symm
        self.barrier_timeout_ms = int(barrier_timeout_ms)
        self.group = dist.group.WORLD if group is None else group
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)

    def allocate(self, nbytes: int) -> _PeerBlock:
        local = self._symm.empty(nbytes, dtype=torch.uint8, device=self.device)
        hdl = self._symm.rendezvous(local, self.group.group_name)

        def views(rank, offset_bytes, shape, dtype):
            item = torch.empty((), dtype=dtype).element_size()
            if rank == self.rank:
                n = int(np.prod(shape)) * item
                return local[offset_bytes:offset_bytes + n].view(dtype).view(shape)
            return hdl.get_buffer(rank, shape, dtype, offset_bytes // item)

        timeout = self.barrier_timeout_ms
        return _PeerBlock(self, local, views, int(hdl.buffer_ptrs_dev), lambda: hdl.barrier(0, timeout))


class EmulatedFabric:
    """G virtual ranks inside ONE process on one GPU (tests): `EmulatedFabric.create(G, device)` returns the G rank
    objects; a peer's block is simply that rank's tensor, the barrier is stream order."""

    def __init__(self, shared, rank, world, device):
        self._shared, self.rank, self.world, self.device = shared, rank, world, torch.device(device)
        self._count = 0

    @classmethod
    def create(cls, world: int, device="cuda"):
        shared = {"blocks": {}, "bases": {}}
        return [cls(shared, r, world, device) for r in range(world)]

    def allocate(self, nbytes: int) -> _PeerBlock:
        idx = self._count
        self._count += 1
        blocks = self._shared["blocks"].setdefault(idx, {})
        local = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
        blocks[self.rank] = local

        def views(rank, offset_bytes, shape, dtype):
            item = torch.empty((), dtype=dtype).element_size()
            n = int(np.prod(shape)) * item
            return blocks[rank][offset_bytes:offset_bytes + n].view(dtype).view(shape)

        block = _PeerBlock(self, local, views, None, lambda: None)
        shared, world, device = self._shared, self.world, self.device

        def bases():            # device array of the G base pointers, built once every rank has allocated
            if idx not in shared["bases"]:
                shared["bases"][idx] = torch.tensor([blocks[r].data_ptr() for r in range(world)], dtype=torch.int64,
                                                    device=device)
            return shared["bases"][idx].data_ptr()

        block._bases_fn = bases
        return block


# ------------------------------------------------------------------------------------------- field solve
class DistributedFieldSolve:
    """Per-rank stages of the distributed spectral field solve (include/fluxgnn.h, fluxgnn_poisson_dist_*).
    Buffers: two flat complex arrays of P*S elements (P = ceil(B/2) complex signals of S cells) that alternate
    as send / receive buffer of the four all-to-alls, plus the FFT scratch."""

    def __init__(self, nx: int, length: float, rank: int, world: int, device, cuda_stages=None):
        if world & (world - 1) or world > 16:
            raise ValueError(f"the distributed field solve needs a power-of-two number of ranks <= 16, got {world}")
        self.nx, self.length, self.rank, self.world = int(nx), float(length), rank, world
        self.S = self.nx // world
        if self.S & (self.S - 1) or self.S < 256:
            raise ValueError(f"the distributed field solve needs power-of-two slabs of >= 256 cells, got {self.S}")
        self.device = torch.device(device)
        self._bufs = {}
        self._stages = cuda_stages          # test hook: CPU stand-ins for the four CUDA stages

    def buffers(self, B: int):
        P = (B + 1) // 2
        hit = self._bufs.get(P)
        if hit is None:
            n = P * self.S * 2
            hit = tuple(torch.empty(n, dtype=torch.float32, device=self.device) for _ in range(3))
            self._bufs[P] = hit
        return hit

    # ---- the four local stages (CUDA; `self._stages` replaces them in the CPU tests) ----
    def pack(self, n_rows: torch.Tensor, z: torch.Tensor):
        """n_rows: [B, S] view (row stride arbitrary, unit column stride) -> z flat [P*S*2] = (n_a - 1, n_b - 1)."""
        if self._stages:
            return self._stages["pack"](n_rows, z)
        B = n_rows.shape[0]
        _lib.check(_lib.lib().fluxgnn_poisson_dist_pack(n_rows.data_ptr(), n_rows.stride(0) if B > 1 else self.S, B, self.S,
                                                        z.data_ptr(), _stream(self.device)), "fluxgnn_poisson_dist_pack")

    def rank_dft(self, src: torch.Tensor, dst: torch.Tensor, inverse: bool):
        if self._stages:
            return self._stages["rank_dft"](self, src, dst, inverse)
        chunk = src.numel() // 2 // self.world
        _lib.check(_lib.lib().fluxgnn_poisson_dist_rank_dft(src.data_ptr(), dst.data_ptr(), self.world, chunk,
                                                            self.rank * chunk, self.S, int(inverse), _stream(self.device)),
                   "fluxgnn_poisson_dist_rank_dft")

    def local(self, y: torch.Tensor, scratch: torch.Tensor):
        if self._stages:
            return self._stages["local"](self, y)
        P = y.numel() // 2 // self.S
        _lib.check(_lib.lib().fluxgnn_poisson_dist_local(y.data_ptr(), scratch.data_ptr(), P, self.S, self.world, self.rank,
                                                         self.length, _stream(self.device)), "fluxgnn_poisson_dist_local")

    def unpack(self, e: torch.Tensor, E_rows: torch.Tensor):
        if self._stages:
            return self._stages["unpack"](e, E_rows)
        B = E_rows.shape[0]
        _lib.check(_lib.lib().fluxgnn_poisson_dist_unpack(e.data_ptr(), E_rows.data_ptr(), E_rows.stride(0) if B > 1 else self.S,
                                                          B, self.S, _stream(self.device)), "fluxgnn_poisson_dist_unpack")

    # ---- one solve on one rank ----
    def solve(self, n_rows: torch.Tensor, E_rows: torch.Tensor, comm):
        """E_rows[B,S] <- field of the global density whose slab is n_rows[B,S] (both may be strided row views)."""
        a, b, scratch = self.buffers(n_rows.shape[0])
        with torch.cuda.device(self.device) if self.device.type == "cuda" else _nullctx():
            self.pack(n_rows, a)
            if self.world > 1:
                comm.all_to_all(b, a)
                self.rank_dft(b, a, False)
                comm.all_to_all(b, a)
            else:
                a, b = b, a
            self.local(b, scratch)
            if self.world > 1:
                comm.all_to_all(a, b)
                self.rank_dft(a, b, True)
                comm.all_to_all(a, b)
            else:
                a, b = b, a
            self.unpack(a, E_rows)


def solve_emulated(solvers, n_rows, E_rows):
    """The same solve for G virtual ranks in one process: every all-to-all becomes chunk copies."""
    G = len(solvers)
    bufs = [s.buffers(n_rows[0].shape[0]) for s in solvers]
    a = [bf[0] for bf in bufs]
    b = [bf[1] for bf in bufs]

    def exchange(dst, src):
        chunk = src[0].numel() // G
        for r in range(G):
            for q in range(G):
                dst[r][q * chunk:(q + 1) * chunk].copy_(src[q][r * chunk:(r + 1) * chunk])

    for r in range(G):
        solvers[r].pack(n_rows[r], a[r])
    if G > 1:
        exchange(b, a)
        for r in range(G):
            solvers[r].rank_dft(b[r], a[r], False)
        exchange(b, a)
    else:
        a, b = b, a
    for r in range(G):
        solvers[r].local(b[r], bufs[r][2])
    if G > 1:
        exchange(a, b)
        for r in range(G):
            solvers[r].rank_dft(a[r], b[r], True)
        exchange(a, b)
    else:
        a, b = b, a
    for r in range(G):
        solvers[r].unpack(a[r], E_rows[r])


class DistributedScanSolve:
    """The field solve as a distributed prefix sum (csrc/scan_poisson.cu, slab form; north_star's "distributed Poisson
    reduction"): every rank sums its slab, ONE all-gather of 48 bytes per IC and rank carries the sums (plus the edge
    densities the neighbours' stencil needs and the certificate sums of the previous field), then every rank
    reconstructs its slab of E.  No transform, no all-to-all.  The certificate (bound on the distance to the spectral
    operator, evaluated for every field) travels one step behind; `first_uncertified` closes the record."""

    MSG_BYTES = 48
    NEVER = 2 ** 31 - 1

    def __init__(self, nx: int, length: float, rank: int, world: int, device, cert_tol: float = 1e-5, cpu_stages=None):
        self.nx, self.length, self.rank, self.world = int(nx), float(length), rank, world
        self.S = self.nx // world
        if self.S % 8 or self.S < 64:
            raise ValueError(f"the distributed scan solve needs slabs of a multiple of 8 cells (>= 64), got {self.S}")
        self.device = torch.device(device)
        self.cert_tol = float(cert_tol)
        self._stages = cpu_stages           # test hook: CPU stand-ins for the CUDA stages
        self._state = {}

    def state(self, B: int):
        st = self._state.get(B)
        if st is None:
            if self._stages:
                st = {"work": None, "msg": torch.zeros(B, 8, dtype=torch.float64), "cert": torch.zeros(B, 2, dtype=torch.float64)}
            else:
                ws = _lib.lib().fluxgnn_scan_slab_workspace_bytes(B, self.S)
                st = {"work": torch.zeros(ws, dtype=torch.uint8, device=self.device),       # zeroed once: carries the certificate sums
                      "msg": torch.empty(B * self.MSG_BYTES, dtype=torch.uint8, device=self.device)}
            st["flag"] = torch.full((1,), self.NEVER, dtype=torch.int32, device=self.device)
            st["step"] = 0
            self._state[B] = st
        return st

    # ---- stages ----
    def sums(self, n_rows: torch.Tensor, st, gather=None):
        """gather = (device pointer table of the ranks' symmetric allocations, offset of the gather buffer in them): the
        message kernel then stores this rank's messages into every rank's gather buffer itself (peer memory)."""
        if self._stages:
            return self._stages["sums"](self, n_rows, st)
        B = n_rows.shape[0]
        if gather is not None:
            _lib.check(_lib.lib().fluxgnn_scan_slab_sums_peer(
                n_rows.data_ptr(), n_rows.stride(0) if B > 1 else self.S, B, self.S, self.rank * self.S, st["work"].data_ptr(),
                st["msg"].data_ptr(), gather[0], gather[1], self.rank, self.world, _stream(self.device)),
                "fluxgnn_scan_slab_sums_peer")
            return
        _lib.check(_lib.lib().fluxgnn_scan_slab_sums(n_rows.data_ptr(), n_rows.stride(0) if B > 1 else self.S, B, self.S,
                                                     self.rank * self.S, st["work"].data_ptr(), st["msg"].data_ptr(),
                                                     _stream(self.device)), "fluxgnn_scan_slab_sums")

    def field(self, n_rows: torch.Tensor, E_rows: torch.Tensor, msg_all: torch.Tensor, st, peers=None):
        """peers = (E rows of the left neighbour's next state, of the right neighbour's, halo): the kernel also stores the
        slab's edge cells of E' into the neighbours' ghost zones (peer memory)."""
        if self._stages:
            return self._stages["field"](self, n_rows, E_rows, msg_all, st)
        B = n_rows.shape[0]
        if peers is not None:
            _lib.check(_lib.lib().fluxgnn_scan_slab_field_peer(
                n_rows.data_ptr(), n_rows.stride(0) if B > 1 else self.S, E_rows.data_ptr(), E_rows.stride(0) if B > 1 else self.S,
                B, self.S, self.rank, self.world, self.length, msg_all.data_ptr(), st["work"].data_ptr(), self.cert_tol,
                st["step"], st["flag"].data_ptr(), peers[0].data_ptr(), peers[1].data_ptr(), peers[2], _stream(self.device)),
                "fluxgnn_scan_slab_field_peer")
            return
        _lib.check(_lib.lib().fluxgnn_scan_slab_field(
            n_rows.data_ptr(), n_rows.stride(0) if B > 1 else self.S, E_rows.data_ptr(), E_rows.stride(0) if B > 1 else self.S,
            B, self.S, self.rank, self.world, self.length, msg_all.data_ptr(), st["work"].data_ptr(), self.cert_tol, st["step"],
            st["flag"].data_ptr(), _stream(self.device)), "fluxgnn_scan_slab_field")

    def certify(self, msg_all: torch.Tensor, st, B: int):
        if self._stages:
            return self._stages["certify"](self, msg_all, st)
        _lib.check(_lib.lib().fluxgnn_scan_slab_certify(B, self.S, self.world, self.length, msg_all.data_ptr(), self.cert_tol,
                                                        st["step"] - 1, st["flag"].data_ptr(), _stream(self.device)),
                   "fluxgnn_scan_slab_certify")

    # ---- one solve on one rank ----
    def solve(self, n_rows: torch.Tensor, E_rows: torch.Tensor, comm):
        """E_rows[B,S] <- field of the global density whose slab is n_rows[B,S] (both may be strided row views)."""
        st = self.state(n_rows.shape[0])
        with torch.cuda.device(self.device) if self.device.type == "cuda" else _nullctx():
            self.sums(n_rows, st)
            msg_all = comm.all_gather(st["msg"])
            self.field(n_rows, E_rows, msg_all, st)
        st["step"] += 1

    def first_uncertified(self, n_rows: torch.Tensor, comm):
        """Index of the first field solve of this solver whose certificate failed (None: all certified).  Every rank
        calls it together; `n_rows` is the current density slab (its sums ride along, they are not used)."""
        st = self.state(n_rows.shape[0])
        if st["step"] == 0:
            return None
        with torch.cuda.device(self.device) if self.device.type == "cuda" else _nullctx():
            self.sums(n_rows, st)
            msg_all = comm.all_gather(st["msg"])
            self.certify(msg_all, st, n_rows.shape[0])
        bad = int(st["flag"].item())
        return None if bad == self.NEVER else bad


def scan_solve_emulated(solvers, n_rows, E_rows):
    """DistributedScanSolve.solve for G virtual ranks in one process: the all-gather becomes a stack."""
    G = len(solvers)
    sts = [solvers[r].state(n_rows[r].shape[0]) for r in range(G)]
    for r in range(G):
        solvers[r].sums(n_rows[r], sts[r])
    msg_all = torch.stack([st["msg"] for st in sts]).contiguous()
    for r in range(G):
        solvers[r].field(n_rows[r], E_rows[r], msg_all, sts[r])
        sts[r]["step"] += 1


def scan_first_uncertified_emulated(solvers, n_rows):
    G = len(solvers)
    sts = [solvers[r].state(n_rows[r].shape[0]) for r in range(G)]
    if sts[0]["step"] == 0:
        return None
    for r in range(G):
        solvers[r].sums(n_rows[r], sts[r])
    msg_all = torch.stack([st["msg"] for st in sts]).contiguous()
    bad = DistributedScanSolve.NEVER
    for r in range(G):
        solvers[r].certify(msg_all, sts[r], n_rows[r].shape[0])
        bad = min(bad, int(sts[r]["flag"].item()))
    return None if bad == DistributedScanSolve.NEVER else bad


class _nullctx:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


# ------------------------------------------------------------------------------------------- slab solvers
class _DomainDecomposedSolver:
    """Shared host logic: extended ping-pong state, halo exchange, field solve, step orchestration."""

    def __init__(self, nx, length, dt, halo, rank, world, device, field_solve, slab_fn=None, field_fn=None,
                 field_stages=None, cert_tol=1e-5, fabric=None):
        self._fabric = fabric                # peer memory (SymmetricMemoryFabric / EmulatedFabric): step_peer(), advance()
        self._peer = {}
        self._graphs = {}
        if fabric is not None:
            if field_solve not in ("auto", "scan"):
                raise ValueError("the peer-memory step uses the distributed prefix-sum field solve: field_solve='scan'")
            field_solve = "scan"
            if (fabric.rank, fabric.world) != (rank, world):
                raise ValueError(f"fabric is rank {fabric.rank} of {fabric.world}, the solver rank {rank} of {world}")
        if nx % world:
            raise ValueError(f"nx={nx} is not divisible by the number of ranks {world}")
        self.nx, self.length, self.dt = int(nx), float(length), float(dt)
        self.rank, self.world, self.device = rank, world, torch.device(device)
        self.owned = self.nx // world
        self.halo = int(halo)
        if self.owned < self.halo:
            raise ValueError(f"a slab of {self.owned} cells is narrower than the halo of {self.halo}")
        if field_solve not in ("auto", "alltoall", "allgather", "scan"):
            raise ValueError("field_solve must be 'auto', 'alltoall', 'allgather' or 'scan'")
        if field_solve == "auto":            # the distributed solve needs power-of-two slabs and rank counts
            ok = not (self.owned & (self.owned - 1)) and self.owned >= 256 and not (world & (world - 1)) and world <= 16
            field_solve = "alltoall" if ok else "allgather"
        self.field_mode = field_solve
        self.grid = PeriodicGrid(self.nx, self.length)
        self.ld = self.owned + 2 * self.halo
        idx = (rank * self.owned - self.halo + np.arange(self.ld)) % self.nx
        self.x_ext = torch.as_tensor(self.grid.x[idx], dtype=torch.float32).to(self.device)   # GLOBAL positions
        self._slab_fn = slab_fn or self._cuda_slab
        self._field_fn = field_fn
        self._baseline = None
        self._ext = {}                       # batch -> [buffer 0, buffer 1]
        self._cur = {}                       # batch -> index of the buffer holding the current state
        self._edges = {}
        self._dist = (DistributedFieldSolve(self.nx, self.length, rank, world, self.device, field_stages)
                      if field_solve == "alltoall" else None)
        # "scan": distributed prefix-sum solve (one all-gather of 48 B per IC and rank); certified per field, see
        # first_uncertified().  Explicit opt-in: a failed certificate is reported, not repaired, on this path.
        self._scan = (DistributedScanSolve(self.nx, self.length, rank, world, self.device, cert_tol, field_stages)
                      if field_solve == "scan" else None)

    # ---- extended state ----
    def _buffers(self, B: int):
        hit = self._ext.get(B)
        if hit is None and self._fabric is not None:
            # one symmetric block: the two extended states, then the gather buffer of the field-solve messages
            rank, world = self.rank, self.world
            ext_bytes = (B * 3 * self.ld * 4 + 255) // 256 * 256
            msg_bytes = B * DistributedScanSolve.MSG_BYTES
            block = self._fabric.allocate(2 * ext_bytes + world * msg_bytes)
            left, right = (rank - 1) % world, (rank + 1) % world
            shape = (B, 3, self.ld)
            hit = [block.view(rank, k * ext_bytes, shape) for k in range(2)]
            for t in hit:
                t.zero_()
            block.barrier()          # a faster neighbour must not push its halo cells before this rank has zeroed its buffers
            self._peer[B] = {"block": block, "msg_off": 2 * ext_bytes, "msg_bytes": msg_bytes, "left": None, "right": None,
                             # the neighbours' views are resolved at the first push (all ranks have allocated by then)
                             "neighbours": lambda: ([block.view(left, k * ext_bytes, shape) for k in range(2)],
                                                    [block.view(right, k * ext_bytes, shape) for k in range(2)]),
                             "msg_all": block.view(rank, 2 * ext_bytes, (world, msg_bytes), torch.uint8)}
            self._ext[B] = hit
            self._cur[B] = 0
            self._edges[B] = None
        if hit is None:
            hit = [torch.zeros(B, 3, self.ld, dtype=torch.float32, device=self.device) for _ in range(2)]
            self._ext[B] = hit
            self._cur[B] = 0
            H = self.halo
            self._edges[B] = (torch.empty(2, B, 3, H, dtype=torch.float32, device=self.device),
                              torch.empty(2, B, 3, H, dtype=torch.float32, device=self.device))
        return hit

    def interior(self, ext: torch.Tensor) -> torch.Tensor:
        return ext[..., self.halo:self.halo + self.owned]

    def _adopt(self, local: torch.Tensor) -> torch.Tensor:
        """The extended buffer that holds `local`: the current one if `local` is its interior view (the value the
        previous step returned), otherwise the state is copied in."""
        if local.dim() != 3 or local.shape[1] != 3 or local.shape[2] != self.owned:
            raise ValueError(f"local state must be [B,3,{self.owned}], got {tuple(local.shape)}")
        B = local.shape[0]
        bufs = self._buffers(B)
        cur = bufs[self._cur[B]]
        mine = self.interior(cur)
        self._fresh = not (local.data_ptr() == mine.data_ptr() and local.stride() == mine.stride())
        if self._fresh:
            mine.copy_(local.to(device=self.device, dtype=torch.float32))
        return cur

    def _fill_ghosts(self, ext: torch.Tensor, comm):
        H, S = self.halo, self.owned
        B = ext.shape[0]
        send, recv = self._edges[B]
        send[0].copy_(ext[..., H:2 * H])             # my first H cells -> left neighbour's right ghosts
        send[1].copy_(ext[..., S:S + H])             # my last H cells  -> right neighbour's left ghosts
        left_ghost, right_ghost = comm.exchange_halos(send[0], send[1])
        ext[..., :H].copy_(left_ghost)
        ext[..., H + S:].copy_(right_ghost)

    # ---- field solve ----
    def field(self, n_full: torch.Tensor) -> torch.Tensor:
        """Global field solve E[B,nx] from the gathered density (src/baseline_solver.py:59-68)."""
        if self._field_fn is not None:
            return self._field_fn(n_full)
        if self._baseline is None:
            self._baseline = BaselineSolver(nx=self.nx, length=self.length, dt=self.dt, device=self.device)
        return self._baseline.solve_poisson(n_full.contiguous())

    def _solve_field(self, nxt: torch.Tensor, comm):
        """E' of the new density (channel 0 of `nxt`'s interior) into channel 2 of `nxt`'s interior."""
        inner = self.interior(nxt)
        if self._scan is not None:
            self._scan.solve(inner[:, 0], inner[:, 2], comm)
            return
        if self._dist is not None:
            self._dist.solve(inner[:, 0], inner[:, 2], comm)
            return
        S = self.owned
        gathered = comm.all_gather(inner[:, 0].contiguous())                    # [world,B,S]
        n_full = gathered.permute(1, 0, 2).reshape(nxt.shape[0], self.nx)
        inner[:, 2].copy_(self.field(n_full)[:, self.rank * S:(self.rank + 1) * S])

    def first_uncertified(self, local: torch.Tensor, comm):
        """field_solve="scan": index of the first step whose field failed its certificate, None if all passed (or if
        another field solve is in use).  Collective: every rank calls it with its current local state."""
        if self._scan is None:
            return None
        return self._scan.first_uncertified(self._adopt(local)[..., self.halo:self.halo + self.owned][:, 0], comm)

    # ---- one step, one process per rank ----
    def step(self, local: torch.Tensor, comm) -> torch.Tensor:
        """local [B,3,owned] -> new local state (the interior VIEW of this solver's extended buffer; pass it back
        to step() and no copy is made).  `comm` provides exchange_halos / all_to_all / all_gather."""
        if self._fabric is not None:
            raise ValueError("this solver was built with fabric=...: its exchange runs over peer memory, call step_peer() / "
                             "advance() (every rank together) instead of step(state, comm)")
        cur = self._adopt(local)
        B = cur.shape[0]
        nxt = self._ext[B][1 - self._cur[B]]
        self._fill_ghosts(cur, comm)
        self._slab_fn(cur, nxt)
        self._solve_field(nxt, comm)
        self._cur[B] = 1 - self._cur[B]
        return self.interior(nxt)

    # ---- one step over peer memory: no NCCL; the exchange is fused into the producing kernels ----
    def _neighbours(self, B: int):
        p = self._peer[B]
        if p["left"] is None:
            p["left"], p["right"] = p["neighbours"]()
        return p

    def _peer_push(self, B: int):
        """My first / last H cells of the CURRENT state (n, u, E) -> the ghost zones of the ring neighbours' current
        buffers (NVLink stores).  Only after a state was copied in from outside; in the steady state the slab and field
        kernels have stored them already."""
        k, p = self._cur[B], self._neighbours(B)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().fluxgnn_peer_halo_push(self._ext[B][k].data_ptr(), p["left"][k].data_ptr(),
                                                         p["right"][k].data_ptr(), B, self.owned, self.halo, 0, 3,
                                                         _stream(self.device)), "fluxgnn_peer_halo_push")

    def _peer_begin(self, local: torch.Tensor) -> int:
        """Adopt `local`; a state copied in from outside also needs its edge cells pushed to the neighbours (every rank
        passes the same kind of state, so the barrier count matches)."""
        B = self._adopt(local).shape[0]
        if self._fresh:
            self._peer_push(B)
            self._peer[B]["block"].barrier()
        return B

    def _peer_compute(self, B: int):
        """Slab kernel: n', u' into the other buffer AND, for the edge cells, into the neighbours' ghost zones of THEIR
        other buffer; then this slab's prefix sums, whose message kernel stores the 48-byte message per IC into every
        rank's gather buffer."""
        k, p = self._cur[B], self._neighbours(B)
        nxt = self._ext[B][1 - k]
        self._slab_fn(self._ext[B][k], nxt, p["left"][1 - k], p["right"][1 - k])
        st = self._scan.state(B)
        block = p["block"]
        bases = block.bases_dev if block.bases_dev is not None else block._bases_fn()
        with torch.cuda.device(self.device):
            self._scan.sums(self.interior(nxt)[:, 0], st, gather=(bases, p["msg_off"]))

    def _peer_field(self, B: int):
        """E' of this slab from the gathered messages, its edge cells also into the neighbours' ghost zones; the new state
        becomes the current one."""
        k, p = self._cur[B], self._neighbours(B)
        inner = self.interior(self._ext[B][1 - k])
        st = self._scan.state(B)
        peers = (self.interior(p["left"][1 - k])[:, 2], self.interior(p["right"][1 - k])[:, 2], self.halo)
        self._scan.field(inner[:, 0], inner[:, 2], p["msg_all"], st, peers)
        st["step"] += 1
        self._cur[B] = 1 - k
        return inner

    def step_peer(self, local: torch.Tensor) -> torch.Tensor:
        """step() over peer memory (the solver was built with fabric=...): slab kernel (+ halo stores of n', u'), slab
        sums, message stores -> barrier -> field kernel (+ halo stores of E') -> barrier.  Every rank calls it together;
        returns the interior view like step()."""
        if self._fabric is None:
            raise ValueError("step_peer needs a solver built with fabric=SymmetricMemoryFabric(...)")
        B = self._peer_begin(local)
        barrier = self._peer[B]["block"].barrier
        self._peer_compute(B)
        barrier()
        out = self._peer_field(B)
        barrier()
        return out

    def advance(self, local: torch.Tensor, n_steps: int, graph: bool = True) -> torch.Tensor:
        """n_steps of step_peer.  graph=True: PAIRS of steps (one per ping-pong buffer) are captured once into a CUDA graph
        -- kernels, peer stores and barriers alike -- and replayed, so a step costs no host work; the first call runs two
        eager steps before it captures.  (The certificate index a replayed step reports is the one baked at capture:
        first_uncertified() then tells THAT a field failed, not exactly which.)"""
        state, done = local, 0
        B = local.shape[0]
        if graph and n_steps >= 2:
            hit = self._graphs.get(B)
            if hit is None and n_steps >= 4:
                state = self.step_peer(self.step_peer(state))
                done = 2
                torch.cuda.synchronize(self.device)
                g = torch.cuda.CUDAGraph()
                st = self._scan.state(B)
                step0, k0 = st["step"], self._cur[B]
                with torch.cuda.graph(g):
                    self.step_peer(self.step_peer(state))
                st["step"] = step0                                # the capture ran nothing
                hit = self._graphs[B] = (g, k0)
            if hit is not None:
                g, k0 = hit
                self._peer_begin(state)                           # a foreign tensor is copied in and its edges pushed
                if self._cur[B] != k0 and done < n_steps:         # the graph starts from buffer k0
                    state = self.step_peer(state)
                    done += 1
                pairs = (n_steps - done) // 2
                for _ in range(pairs):
                    g.replay()
                if pairs:
                    self._scan.state(B)["step"] += 2 * pairs
                    done += 2 * pairs
                    state = self.interior(self._ext[B][self._cur[B]])
        while done < n_steps:
            state = self.step_peer(state)
            done += 1
        return state

    def time_shares(self, local: torch.Tensor, comm, reps: int = 3):
        """Device time of a full step against the slab kernel alone (all ranks call this together):
        {"step_ms", "slab_kernel_ms", "exchange_and_solve_share"}."""
        cur = self._adopt(local)
        B = cur.shape[0]
        nxt = self._ext[B][1 - self._cur[B]]
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        torch.cuda.synchronize(self.device)
        ev[0].record()
        for _ in range(reps):
            self._slab_fn(cur, nxt)
        ev[1].record()
        state = local
        ev[2].record()
        for _ in range(reps):
            state = self.step(state, comm)
        ev[3].record()
        torch.cuda.synchronize(self.device)
        slab, full = ev[0].elapsed_time(ev[1]) / reps, ev[2].elapsed_time(ev[3]) / reps
        return {"step_ms": full, "slab_kernel_ms": slab, "exchange_and_solve_share": max(0.0, 1.0 - slab / full)}


class DomainDecomposedHybridSolver(_DomainDecomposedSolver):
    """The hybrid step on one slab of a grid of `nx` cells split over `world` ranks."""

    def __init__(self, model, nx, length=2 * np.pi, dt=5e-3, graph_radius=1, rank=0, world=1, device="cuda",
                 precision="fp32", slab_fn=None, field_fn=None, field_solve="auto", field_stages=None, fabric=None):
        self.model = model
        self.radius = int(graph_radius)
        if precision != "fp32" and precision not in _lib.TC_PRECISIONS:
            raise ValueError(f"precision must be 'fp32' or one of {sorted(_lib.TC_PRECISIONS)}, got {precision!r}")
        self.precision = precision
        super().__init__(nx, length, dt, model.num_layers * self.radius + 1, rank, world, device, field_solve,
                         slab_fn, field_fn, field_stages, fabric=fabric)

    def _cuda_slab(self, ext: torch.Tensor, nxt: torch.Tensor, left=None, right=None):
        """left / right: the ring neighbours' next extended states (peer memory) -- the kernel then stores the edge cells of
        n', u' into their ghost zones too."""
        tensor_path = self.precision != "fp32"
        packed = self.model.packed_weights(_lib.weight_layout(self.precision))
        B = ext.shape[0]
        with torch.cuda.device(self.device):
            dx = self.length / self.nx
            head = (packed.data_ptr(), self.model.num_layers, _lib.TC_PRECISIONS[self.precision] if tensor_path else 0,
                    ext.data_ptr(), self.x_ext.data_ptr(), nxt.data_ptr(), self.ld, self.halo, B, self.owned, self.halo,
                    self.radius, float(np.float32(self.dt / dx)), float(np.float32(self.dt)))
            if left is not None:
                _lib.check(_lib.lib().fluxgnn_hybrid_slab_step_peer(*head, left.data_ptr(), right.data_ptr(),
                                                                    _stream(self.device)), "fluxgnn_hybrid_slab_step_peer")
            else:
                _lib.check(_lib.lib().fluxgnn_hybrid_slab_step_ld(*head, _stream(self.device)), "fluxgnn_hybrid_slab_step_ld")

    def advance_slab(self, ext: torch.Tensor) -> torch.Tensor:
        """ext [B,3,owned+2*halo] (ghosts included) -> [B,3,owned] with n', u' (E' not yet)."""
        if ext.shape[1:] != (3, self.ld):
            raise ValueError(f"ext must be [B,3,{self.ld}], got {tuple(ext.shape)}")
        nxt = torch.zeros_like(ext)
        self._slab_fn(ext.to(torch.float32).contiguous(), nxt)
        return self.interior(nxt).contiguous()


class DomainDecomposedBaselineSolver(_DomainDecomposedSolver):
    """The classical step (src/baseline_solver.py:80-101) on one slab: halo exchange of 4 cells (1 is read),
    fluxgnn_baseline_slab_step, distributed field solve."""

    def __init__(self, nx, length=2 * np.pi, dt=5e-3, nu=1e-3, rank=0, world=1, device="cuda", slab_fn=None,
                 field_fn=None, field_solve="auto", field_stages=None, fabric=None):
        self.nu = float(nu)
        super().__init__(nx, length, dt, 4, rank, world, device, field_solve, slab_fn, field_fn, field_stages,
                         fabric=fabric)

    def _cuda_slab(self, ext: torch.Tensor, nxt: torch.Tensor, left=None, right=None):
        B = ext.shape[0]
        dx = self.length / self.nx
        with torch.cuda.device(self.device):
            head = (ext.data_ptr(), nxt.data_ptr(), self.ld, self.halo, None, B, self.owned, self.halo,
                    float(np.float32(self.dt / dx)), float(np.float32(self.dt)), float(np.float32(self.nu)),
                    float(np.float32(dx ** 2)))
            if left is not None:
                _lib.check(_lib.lib().fluxgnn_baseline_slab_step_peer(*head, left.data_ptr(), right.data_ptr(),
                                                                      _stream(self.device)), "fluxgnn_baseline_slab_step_peer")
            else:
                _lib.check(_lib.lib().fluxgnn_baseline_slab_step(*head, _stream(self.device)), "fluxgnn_baseline_slab_step")


def split_slabs(state: torch.Tensor, world: int):
    """[B,3,nx] -> list of `world` contiguous slabs [B,3,nx/world]."""
    return [s.contiguous() for s in state.chunk(world, dim=-1)]


def step_emulated(solvers, locals_):
    """One decomposed step of G virtual ranks inside this process (solvers[r] built with rank=r,
    world=G): the halo exchange and the collectives become tensor copies.  For single-GPU tests."""
    G = len(solvers)
    H, S = solvers[0].halo, solvers[0].owned
    curs = [solvers[r]._adopt(locals_[r]) for r in range(G)]
    B = curs[0].shape[0]
    nxts = [solvers[r]._ext[B][1 - solvers[r]._cur[B]] for r in range(G)]
    for r in range(G):
        curs[r][..., :H].copy_(curs[(r - 1) % G][..., S:S + H])
        curs[r][..., H + S:].copy_(curs[(r + 1) % G][..., H:2 * H])
    for r in range(G):
        solvers[r]._slab_fn(curs[r], nxts[r])
    inner = [solvers[r].interior(nxts[r]) for r in range(G)]
    if solvers[0]._scan is not None:
        scan_solve_emulated([s._scan for s in solvers], [i[:, 0] for i in inner], [i[:, 2] for i in inner])
    elif solvers[0]._dist is not None:
        solve_emulated([s._dist for s in solvers], [i[:, 0] for i in inner], [i[:, 2] for i in inner])
    else:
        n_full = torch.cat([i[:, 0] for i in inner], dim=-1)
        for r in range(G):
            inner[r][:, 2].copy_(solvers[r].field(n_full)[:, r * S:(r + 1) * S])
    for r in range(G):
        solvers[r]._cur[B] = 1 - solvers[r]._cur[B]
    return inner


def step_peer_emulated(solvers, locals_):
    """step_peer for G virtual ranks in one process (solvers built with the objects of EmulatedFabric.create(G)): the
    three phases run rank by rank, stream order stands in for the barriers.  Exercises the real peer kernels."""
    G = len(solvers)
    B = [solvers[r]._adopt(locals_[r]).shape[0] for r in range(G)][0]
    for r in range(G):
        if solvers[r]._fresh:
            solvers[r]._peer_push(B)
    for r in range(G):
        solvers[r]._peer_compute(B)
    return [solvers[r]._peer_field(B) for r in range(G)]
