"""BaselineSolver, drop-in for src/baseline_solver.py of the reference, on sm_100a.

Same constructor, attributes (`nx, length, dx, dt, t_end, x, n0, nu, k`) and
methods; numpy [3, nx] float32 in -> numpy out as in the reference.  Every
method additionally accepts a batch: a numpy [B, 3, nx] array, or a CUDA
tensor [B, 3, nx] that is advanced without leaving the device.

Arithmetic runs in libfluxgnn.so (no CPU fallback); only the random initial
condition generator, which is an input generator rather than part of the
step, stays on the host so that its numpy RandomState draw order is kept
(src/baseline_solver.py:29-57).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib
from .grid import PeriodicGrid


_STAGING_MIN_BYTES = 1 << 18     # below this a pageable copy is as fast as staging through pinned memory
_staging = {}                    # (device, nbytes, role) -> pinned uint8 buffer, reused across calls
_staging_events = {}             # data_ptr of an input staging buffer -> event after its last host-to-device copy


def _pinned(device, nbytes: int, role: str) -> torch.Tensor:
    key = (str(device), nbytes, role)
    buf = _staging.get(key)
    if buf is None:
        if len(_staging) >= 16:              # a handful of shapes at most; do not hoard pinned memory
            torch.cuda.synchronize()
            _staging.clear()
            _staging_events.clear()
        buf = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        _staging[key] = buf
    return buf


def _as_device_batch(state, device):
    """-> (tensor [B,3,nx] float32 contiguous on device, kind) with kind in
    {'np1','npB','t1','tB'} describing how to hand the result back.  Large numpy inputs travel through a reused
    pinned staging buffer (one host memcpy + an asynchronous DMA instead of a pageable copy)."""
    if isinstance(state, np.ndarray):
        arr = np.ascontiguousarray(state, dtype=np.float32)
        kind = "np1" if arr.ndim == 2 else "npB"
        arr3 = arr if arr.ndim == 3 else arr[None]
        dev = torch.device(device)
        if dev.type == "cuda" and arr3.nbytes >= _STAGING_MIN_BYTES:
            raw = _pinned(dev, arr3.nbytes, "in")
            pin = raw.view(torch.float32).view(arr3.shape)
            done = _staging_events.get(raw.data_ptr())
            if done is not None:
                done.synchronize()          # the previous call's DMA out of this buffer must be over before it is overwritten
            pin.numpy()[...] = arr3
            out = pin.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            _staging_events[raw.data_ptr()] = ev
            return out, kind
        return torch.from_numpy(arr3).to(device, non_blocking=False), kind
    if not torch.is_tensor(state):
        raise TypeError("state must be a numpy array or a torch tensor")
    kind = "t1" if state.dim() == 2 else "tB"
    t = state if state.dim() == 3 else state.unsqueeze(0)
    return t.to(device=device, dtype=torch.float32).contiguous(), kind


def _hand_back(t: torch.Tensor, kind: str, lead: int = 0):
    """Undo _as_device_batch; `lead` = number of leading (time) axes in front of [B,3,nx]."""
    if kind in ("np1", "t1"):
        t = t.select(lead, 0)
    if not kind.startswith("np"):
        return t
    nbytes = t.numel() * t.element_size()
    if t.is_cuda and t.dtype == torch.float32 and nbytes >= _STAGING_MIN_BYTES:
        pin = _pinned(t.device, nbytes, "out").view(torch.float32).view(t.shape)
        pin.copy_(t, non_blocking=True)
        torch.cuda.current_stream(t.device).synchronize()
        return pin.numpy().copy()           # a fresh array, as the reference returns
    return t.cpu().numpy()


class BaselineSolver:
    """1D fluid-Poisson model with viscosity (upwind fluxes, forward Euler, spectral field solve)."""

    FIELD_SOLVES = ("auto", "spectral", "scan")

    def __init__(self, nx=64, length=2 * math.pi, dt=5e-3, t_end=1.0, nu=1e-3, *, device="cuda",
                 field_solve="auto", cert_tol=1e-5):
        """`field_solve` (keyword-only, not in the reference): how E = Re ifft(i fft(n-1)/k) is evaluated on long grids.
        "spectral": the FFT kernels, always.  "scan": the certified prefix-sum solve (csrc/scan_poisson.cu; grids of
        at least 4096 cells, nx % 8 == 0), raising if a field's deviation bound exceeds cert_tol * max|E|.  "auto"
        (default): the scan solve where it applies; an IC whose certificate fails is repeated with the FFT solve, so the
        result is always within cert_tol * max|E| of the spectral operator.  `last_field_solve` tells what ran ("scan",
        "spectral" or "scan+spectral"), `last_uncertified_ics` which ICs were repeated."""
        if field_solve not in self.FIELD_SOLVES:
            raise ValueError(f"field_solve must be one of {self.FIELD_SOLVES}, got {field_solve!r}")
        self.field_solve, self.cert_tol = field_solve, float(cert_tol)
        self.last_field_solve, self.last_uncertified_step, self.last_uncertified_ics = None, None, []
        self.grid = PeriodicGrid(nx, length)
        self.nx, self.length, self.dx = self.grid.nx, self.grid.length, self.grid.dx
        self.dt, self.t_end, self.nu = dt, t_end, nu
        self.x, self.k = self.grid.x, self.grid.k
        self.n0 = 1.0
        self.device = torch.device(device)
        if self.dt / self.dx > 0.5:
            print("Warning: dt/dx may be large; consider reducing dt for stability.")

    # ---- scalars exactly as numpy's weak python floats act on float32 arrays ----
    @property
    def _c(self):
        return float(np.float32(self.dt / self.dx))

    # ------------------------------------------------------------------ inputs
    def initial_condition(self, seed=None):
        """Random-mode initial condition [3, nx] float32; draw order as the reference:
        randint(3,6) density modes, each (randint(1,6), rand, rand); two velocity
        modes likewise; then randn(nx) noise (src/baseline_solver.py:29-57)."""
        rng = np.random.RandomState(seed)
        x = self.x

        def mode(lo, span, trig):
            k_mode = rng.randint(1, 6)
            amp = lo + span * rng.rand()
            phase = 2 * np.pi * rng.rand()
            return amp * trig(k_mode * x + phase).astype(np.float32)

        n = np.full(self.nx, self.n0, dtype=np.float32)
        for _ in range(rng.randint(3, 6)):
            n += mode(0.15, 0.15, np.sin)
        u = np.zeros(self.nx, dtype=np.float32)
        for _ in range(2):
            u += mode(0.1, 0.1, np.cos)
        u += 0.05 * rng.randn(self.nx).astype(np.float32)
        E = self.solve_poisson(n)
        return np.stack([n, u, np.asarray(E, dtype=np.float32)], axis=0).astype(np.float32)

    # ------------------------------------------------------------------ field solve
    def solve_poisson(self, n):
        """E = Re ifft(i fft(n - n0)/k) (src/baseline_solver.py:59-68); n is [nx] or [B, nx]."""
        is_np = isinstance(n, np.ndarray)
        t = torch.from_numpy(np.ascontiguousarray(n, dtype=np.float32)) if is_np else n
        single = t.dim() == 1
        t = (t.unsqueeze(0) if single else t).to(device=self.device, dtype=torch.float32).contiguous()
        _, gtab = self.grid.tables(self.device)
        with torch.cuda.device(self.device):
            out = torch.empty_like(t)
            ws_bytes = _lib.lib().fluxgnn_poisson_workspace_bytes(t.shape[0], self.nx)
            work = torch.empty(ws_bytes // 4, dtype=torch.float32, device=self.device) if ws_bytes else None
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(_lib.lib().fluxgnn_poisson_spectral(
                t.data_ptr(), self.nx, out.data_ptr(), self.nx, gtab.data_ptr() if gtab is not None else None,
                t.shape[0], self.nx, self.length, work.data_ptr() if work is not None else None, stream),
                "fluxgnn_poisson_spectral")
        out = out[0] if single else out
        return out.cpu().numpy() if is_np else out

    # ---- small host helpers kept for API compatibility (src/baseline_solver.py:70-78) ----
    def compute_flux_n(self, n, u):
        return (n * u).astype(np.float32)

    def compute_flux_u(self, u):
        return (0.5 * u * u).astype(np.float32)

    def laplacian_u(self, u):
        return (np.roll(u, -1) - 2 * u + np.roll(u, 1)) / (self.dx ** 2)

    # ------------------------------------------------------------------ stepping
    def rollout(self, state: torch.Tensor, n_steps: int, record_every: int = 0, record_flux: bool = False,
                field_solve: str | None = None):
        """Device-resident rollout of state [B,3,nx] (CUDA float32).
        Returns (final [B,3,nx], traj [n_steps//record_every, B,3,nx] | None, flux_n [n_steps,B,nx] | None).
        `field_solve` overrides the constructor's choice for this call."""
        if not torch.is_tensor(state) or state.dim() != 3 or state.shape[1] != 3 or state.shape[2] != self.nx:
            raise ValueError(f"state must be a tensor [B,3,{self.nx}], got {tuple(getattr(state, 'shape', ()))}")
        dev = self.device
        if dev.type != "cuda":
            raise _lib.FluxGNNError(f"BaselineSolver(device={dev}): the step runs in libfluxgnn.so on a CUDA device "
                                    "(there is no CPU fallback)")
        mode = self.field_solve if field_solve is None else field_solve
        if mode not in self.FIELD_SOLVES:
            raise ValueError(f"field_solve must be one of {self.FIELD_SOLVES}, got {mode!r}")
        state = state.to(device=dev, dtype=torch.float32).contiguous()
        B, _, nx = state.shape
        if n_steps < 0:
            raise ValueError(f"n_steps must be >= 0, got {n_steps}")
        if n_steps == 0:                      # the reference's loops simply do not run (src/baseline_solver.py:106-116)
            return (state.clone(), state.new_empty((0, B, 3, nx)) if record_every else None,
                    state.new_empty((0, B, nx)) if record_flux else None)
        L = _lib.lib()
        with torch.cuda.device(dev):
            traj = (torch.empty(n_steps // record_every, B, 3, nx, dtype=torch.float32, device=dev)
                    if record_every else None)
            flux = torch.empty(n_steps, B, nx, dtype=torch.float32, device=dev) if record_flux else None
            scan_ok = bool(L.fluxgnn_baseline_scan_supported(B, nx))
            if mode == "scan" and not scan_ok:
                raise _lib.FluxGNNError(f"field_solve='scan' needs nx >= 4096 and nx % 8 == 0 (B={B}, nx={nx})")
            if mode == "spectral" or not scan_ok:
                out = self._rollout_spectral(state, n_steps, record_every, traj, flux)
                self.last_field_solve, self.last_uncertified_step, self.last_uncertified_ics = "spectral", None, []
                return out, traj, flux
            # "auto" / "scan".  Certificates are per IC: an IC whose certificate fails is repeated with the FFT solve,
            # the others keep the scan result.  The rounding noise of n -- and with it the (conservative) bound -- can
            # only grow with the step count, so a long "auto" rollout is cut into chunks and an IC that failed once
            # stays with the FFT solve: at most one chunk of its work is wasted.
            chunk = n_steps
            if mode == "auto" and n_steps > self.AUTO_CHUNK:
                chunk = self.AUTO_CHUNK if not record_every else max(record_every, self.AUTO_CHUNK // record_every * record_every)
            cur, done = state, 0
            fft_ics = torch.zeros(B, dtype=torch.bool)                   # (host) ICs that left the scan path
            self.last_uncertified_step = None
            while done < n_steps:
                k = min(chunk, n_steps - done)
                tv = traj[done // record_every:(done + k) // record_every] if traj is not None else None
                fv = flux[done:done + k] if flux is not None else None
                if bool(fft_ics.all()):
                    out = self._rollout_spectral(cur, k, record_every, tv, fv)
                else:
                    out, first_bad = self._rollout_scan(cur, k, record_every, tv, fv)          # [B] int32, INT_MAX = certified
                    failed = first_bad != 2 ** 31 - 1
                    if bool(failed.any()):
                        step_bad = done + int(first_bad.min())
                        if self.last_uncertified_step is None:
                            self.last_uncertified_step = step_bad
                        if mode == "scan":
                            raise _lib.FluxGNNError(
                                f"field_solve='scan': the field of step {step_bad} (IC {int(first_bad.argmin())}) is not "
                                f"certified to cert_tol={self.cert_tol:g} (rough density, weak field, short grid, or rounding "
                                "noise accumulated over many steps); use field_solve='auto' or 'spectral'")
                    fft_ics |= failed
                    if bool(fft_ics.any()):
                        idx = fft_ics.nonzero().flatten().to(dev)
                        sub_t = tv.new_empty((tv.shape[0], idx.numel(), 3, nx)) if tv is not None else None
                        sub_f = fv.new_empty((k, idx.numel(), nx)) if fv is not None else None
                        sub = self._rollout_spectral(cur.index_select(0, idx).contiguous(), k, record_every, sub_t, sub_f)
                        out.index_copy_(0, idx, sub)
                        if tv is not None:
                            tv.index_copy_(1, idx, sub_t)
                        if fv is not None:
                            fv.index_copy_(1, idx, sub_f)
                cur, done = out, done + k
            n_fft = int(fft_ics.sum())
            self.last_uncertified_ics = fft_ics.nonzero().flatten().tolist()
            self.last_field_solve = "scan" if n_fft == 0 else ("spectral" if n_fft == B else "scan+spectral")
        return cur, traj, flux

    AUTO_CHUNK = 128

    def _scalars(self):
        return (self._c, float(np.float32(self.dt)), float(np.float32(self.nu)), float(np.float32(self.dx ** 2)))

    def _rollout_scan(self, state, n_steps, record_every, traj, flux):
        """fluxgnn_baseline_rollout_scan -> (final state, host int32[B]: first uncertified step per IC, INT_MAX = none)."""
        L, dev = _lib.lib(), self.device
        B, _, nx = state.shape
        out = torch.empty_like(state)
        work = torch.empty(L.fluxgnn_baseline_scan_workspace_bytes(B, nx) // 4, dtype=torch.float32, device=dev)
        flag = torch.empty(B, dtype=torch.int32, device=dev)
        _lib.check(L.fluxgnn_baseline_rollout_scan(
            state.data_ptr(), out.data_ptr(), B, nx, self.length, *self._scalars(), n_steps, max(record_every, 1),
            traj.data_ptr() if traj is not None else None, flux.data_ptr() if flux is not None else None,
            self.cert_tol, work.data_ptr(), flag.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
            "fluxgnn_baseline_rollout_scan")
        return out, flag.cpu()                        # read-back of B ints: was every reconstructed field certified?

    def _rollout_spectral(self, state, n_steps, record_every, traj, flux):
        L, dev = _lib.lib(), self.device
        B, _, nx = state.shape
        out = torch.empty_like(state)
        _, gtab = self.grid.tables(dev)
        ws_bytes = L.fluxgnn_baseline_workspace_bytes(B, nx)
        work = torch.empty(ws_bytes // 4, dtype=torch.float32, device=dev) if ws_bytes else None
        _lib.check(L.fluxgnn_baseline_rollout(
            state.data_ptr(), out.data_ptr(), gtab.data_ptr() if gtab is not None else None, B, nx, self.length,
            *self._scalars(), n_steps, max(record_every, 1), traj.data_ptr() if traj is not None else None,
            flux.data_ptr() if flux is not None else None,
            work.data_ptr() if work is not None else None, torch.cuda.current_stream(dev).cuda_stream),
            "fluxgnn_baseline_rollout")
        return out

    def step(self, state, return_flux=False):
        """One step (src/baseline_solver.py:80-101); with return_flux also the continuity flux F_n."""
        dev_state, kind = _as_device_batch(state, self.device)
        out, _, flux = self.rollout(dev_state, 1, record_flux=return_flux)
        new = _hand_back(out, kind)
        if return_flux:
            return new, _hand_back(flux[0], kind)
        return new

    def run(self, state0, n_steps=10, record_flux=True):
        """[T+1,3,nx] states including state0 and [T,nx] fluxes (src/baseline_solver.py:103-118)."""
        dev_state, kind = _as_device_batch(state0, self.device)
        _, traj, flux = self.rollout(dev_state, n_steps, record_every=1, record_flux=record_flux)
        states = torch.cat([dev_state.unsqueeze(0), traj], dim=0)
        return _hand_back(states, kind, lead=1), (_hand_back(flux, kind, lead=1) if record_flux else None)
