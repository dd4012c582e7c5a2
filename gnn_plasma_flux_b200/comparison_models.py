"""The reference's two comparison models on sm_100a.

    PureGNN(input_dim=4, hidden_dim=64, num_layers=3)     scripts/training/train_pure_gnn.py:35-76
        .forward(node_features[nx,4], edge_index[2,2nx]) -> delta_state [nx,3]
        .rollout(state[B,3,nx], x[nx], n_steps)            the loop of scripts/evaluation/benchmark_timing.py:129-143
    PINN(input_dim=3*64, hidden_dim=256, num_layers=4)    scripts/training/train_pinn.py:36-61
        .forward(state[...,3,nx]) -> state + delta         (benchmark_timing.py:186-189 applies it step by step)

Same constructor arguments, parameter names (`state_dict` keys) and return types as the reference
classes, so their checkpoints load unchanged.  Inference only; the arithmetic runs in
libfluxgnn.so (csrc/comparison_kernels.cu) -- there is no CPU or eager fallback.  PureGNN accepts
the nearest-neighbour ring of `build_chain_graph` (the only graph the reference ever feeds it),
hidden_dim 64 or 128 and grids of up to 128 cells.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _lib
from .graph_constructor import ring_edge_index


def _f32(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


class PureGNN(nn.Module):
    def __init__(self, input_dim=4, hidden_dim=64, num_layers=3):
        super().__init__()
        self.input_dim, self.hidden_dim, self.num_layers = input_dim, hidden_dim, num_layers
        self.input_mlp = nn.Sequential(nn.Linear(input_dim, hidden_dim), nn.Tanh())
        self.update_mlps = nn.ModuleList([nn.Sequential(nn.Linear(hidden_dim * 2, hidden_dim), nn.Tanh())
                                          for _ in range(num_layers)])
        self.output_mlp = nn.Sequential(nn.Linear(hidden_dim, hidden_dim), nn.Tanh(), nn.Linear(hidden_dim, 3))
        self._packed = None

    def invalidate_packed(self):
        """Repack on the next call (needed after edits through `p.data`, which bump no version counter)."""
        self._packed = None
        return self

    def load_state_dict(self, *args, **kwargs):
        self._packed = None
        return super().load_state_dict(*args, **kwargs)

    def train(self, mode: bool = True):
        self._packed = None
        return super().train(mode)

    def _apply(self, fn, *args, **kwargs):
        self._packed = None
        return super()._apply(fn, *args, **kwargs)

    def _packed_weights(self) -> torch.Tensor:
        if self.input_dim != 4 or self.hidden_dim not in (64, 128) or not 1 <= self.num_layers <= 8:
            raise NotImplementedError("the sm_100a PureGNN kernel supports input_dim=4, hidden_dim in {64, 128}, 1..8 layers")
        params = list(self.parameters())
        dev = params[0].device
        if dev.type != "cuda":
            raise _lib.FluxGNNError("PureGNN parameters are on %s: the forward pass needs a CUDA device" % dev)
        key = tuple((p.data_ptr(), p._version) for p in params)
        if self._packed is None or self._packed[0] != key:
            with torch.cuda.device(dev), torch.no_grad():
                parts = [_f32(self.input_mlp[0].weight), _f32(self.input_mlp[0].bias),
                         torch.stack([_f32(m[0].weight) for m in self.update_mlps]).contiguous(),
                         torch.stack([_f32(m[0].bias) for m in self.update_mlps]).contiguous(),
                         _f32(self.output_mlp[0].weight), _f32(self.output_mlp[0].bias),
                         _f32(self.output_mlp[2].weight), _f32(self.output_mlp[2].bias)]
                L = _lib.lib()
                packed = torch.empty(L.fluxgnn_pure_gnn_packed_bytes(self.hidden_dim, self.num_layers) // 4,
                                     dtype=torch.float32, device=dev)
                _lib.check(L.fluxgnn_pure_gnn_pack(*[t.data_ptr() for t in parts], self.hidden_dim, self.num_layers,
                                                   packed.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
                           "fluxgnn_pure_gnn_pack")
            self._packed = (key, packed)
        return self._packed[1]

    @torch.no_grad()
    def rollout(self, state: torch.Tensor, x: torch.Tensor, n_steps: int = 1) -> torch.Tensor:
        """state [B,3,nx] -> state after n_steps of  state += PureGNN([n,u,E,x], ring)  (one kernel launch)."""
        packed = self._packed_weights()
        dev = packed.device
        state = state.to(device=dev, dtype=torch.float32).contiguous()
        x = x.to(device=dev, dtype=torch.float32).contiguous()
        B, _, nx = state.shape
        with torch.cuda.device(dev):
            out = torch.empty_like(state)
            _lib.check(_lib.lib().fluxgnn_pure_gnn_rollout(
                packed.data_ptr(), self.hidden_dim, self.num_layers, state.data_ptr(), out.data_ptr(), x.data_ptr(),
                B, nx, n_steps, torch.cuda.current_stream(dev).cuda_stream), "fluxgnn_pure_gnn_rollout")
        return out

    @torch.no_grad()
    def forward(self, node_features: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
        nx = node_features.shape[0]
        if node_features.dim() != 2 or node_features.shape[1] != 4:
            raise ValueError(f"node_features must be [nx,4], got {tuple(node_features.shape)}")
        ring = ring_edge_index(nx, 1, device=edge_index.device)
        if tuple(edge_index.shape) != tuple(ring.shape) or not torch.equal(edge_index.to(ring.dtype), ring):
            raise NotImplementedError("PureGNN.forward runs on the nearest-neighbour ring of build_chain_graph only")
        packed = self._packed_weights()
        dev = packed.device
        state = node_features[:, :3].t().unsqueeze(0).to(device=dev, dtype=torch.float32).contiguous()
        x = node_features[:, 3].to(device=dev, dtype=torch.float32).contiguous()
        with torch.cuda.device(dev):
            delta = torch.empty_like(state)           # the kernel emits the model output itself, not (state+delta)-state
            _lib.check(_lib.lib().fluxgnn_pure_gnn_delta(
                packed.data_ptr(), self.hidden_dim, self.num_layers, state.data_ptr(), delta.data_ptr(), x.data_ptr(),
                1, nx, torch.cuda.current_stream(dev).cuda_stream), "fluxgnn_pure_gnn_delta")
        return delta[0].t().contiguous().to(node_features.device)


class PINN(nn.Module):
    def __init__(self, input_dim=3 * 64, hidden_dim=256, num_layers=4):
        super().__init__()
        layers = [nn.Linear(input_dim, hidden_dim), nn.Tanh()]
        for _ in range(num_layers - 2):
            layers += [nn.Linear(hidden_dim, hidden_dim), nn.Tanh()]
        layers.append(nn.Linear(hidden_dim, input_dim))
        self.net = nn.Sequential(*layers)

    @torch.no_grad()
    def forward(self, state: torch.Tensor) -> torch.Tensor:
        linears = [m for m in self.net if isinstance(m, nn.Linear)]
        dev = linears[0].weight.device
        if dev.type != "cuda":
            raise _lib.FluxGNNError("PINN parameters are on %s: the forward pass needs a CUDA device" % dev)
        if state.dim() < 2 or state.shape[-2] * state.shape[-1] != linears[0].in_features:
            raise ValueError(f"PINN was built for {linears[0].in_features} inputs (3*nx); got a state of shape "
                             f"{tuple(state.shape)}")
        batch_shape = state.shape[:-2]
        flat = state.to(device=dev, dtype=torch.float32).reshape(-1, state.shape[-2] * state.shape[-1]).contiguous()
        rows = flat.shape[0]
        L = _lib.lib()
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            cur = flat
            for i, lin in enumerate(linears):
                last = i == len(linears) - 1
                out = torch.empty(rows, lin.out_features, dtype=torch.float32, device=dev)
                w, b = _f32(lin.weight), _f32(lin.bias)
                _lib.check(L.fluxgnn_dense_layer(cur.data_ptr(), w.data_ptr(), b.data_ptr(),
                                                 flat.data_ptr() if last else None, out.data_ptr(), rows,
                                                 lin.in_features, lin.out_features, 0 if last else 1, stream),
                           "fluxgnn_dense_layer")
                cur = out
        return cur.reshape(*batch_shape, state.shape[-2], state.shape[-1])
