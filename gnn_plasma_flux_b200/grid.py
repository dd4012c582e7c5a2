"""Periodic cell-centred grid and its device-side tables.

Restates the constructor arithmetic of src/baseline_solver.py:14-27 (dx, x, k)
and owns the two device arrays every kernel call needs: the float32 cell
centres (node feature 4, src/graph_constructor.py:30) and the float64
circular-convolution table of the spectral field solve.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib


class PeriodicGrid:
    def __init__(self, nx: int = 64, length: float = 2 * math.pi):
        self.nx = int(nx)
        self.length = float(length)
        self.dx = self.length / self.nx
        self.x = np.linspace(0.5 * self.dx, self.length - 0.5 * self.dx, self.nx)
        self.k = 2.0 * np.pi * np.fft.fftfreq(self.nx, d=self.dx)
        self._dev = {}

    def tables(self, device):
        """(x float32 [nx], gtab float64 [nx] | None) on `device`, built once per device.
        gtab is only built where a kernel reads it: whole-IC tiles (nx <= 128) and grids
        whose standalone field solve is the direct convolution rather than the FFT."""
        device = torch.device(device)
        key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
        hit = self._dev.get(key)
        if hit is None:
            with torch.cuda.device(device):
                x_dev = torch.as_tensor(self.x, dtype=torch.float32).to(device)
                gtab = None
                if self.nx <= 128 or _lib.lib().fluxgnn_poisson_uses_table(self.nx):
                    gtab = torch.empty(self.nx, dtype=torch.float64, device=device)
                    stream = torch.cuda.current_stream(device).cuda_stream
                    _lib.check(_lib.lib().fluxgnn_poisson_table(self.nx, self.length, gtab.data_ptr(), stream),
                               "fluxgnn_poisson_table")
                stream = torch.cuda.current_stream(device)
                ready = torch.cuda.Event()
                ready.record(stream)
            hit = (x_dev, gtab, ready, stream.cuda_stream)
            self._dev[key] = hit
        # a consumer on another stream waits for the table kernel / H2D copy of the first call
        cur = torch.cuda.current_stream(device)
        if cur.cuda_stream != hit[3] and not torch.cuda.is_current_stream_capturing():
            cur.wait_event(hit[2])
        return hit[0], hit[1]
