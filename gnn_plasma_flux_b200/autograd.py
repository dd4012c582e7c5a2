"""Autograd support for FluxGNN.forward on the ring.

The reference trains through `model(node_features, edge_index)` with PyTorch autograd
(scripts/training/train_ablation.py:128-206).  Here the forward pass is the fused sm_100a kernel
with activations saved, and the backward pass is hand-written CUDA (csrc/train_kernels.cu) behind
`fluxgnn_backward_ring`; PyTorch only carries the tensors and the graph edge.
"""
from __future__ import annotations

import torch

from . import _lib


class RingFluxes(torch.autograd.Function):
    """flux_edges[B, 2*hops*nx] = FluxGNN(state[B,3,nx], x[nx]) on the radius-r ring, differentiable
    w.r.t. `state` and the parameters (passed explicitly so that autograd tracks them)."""

    @staticmethod
    def forward(ctx, model, state, x, radius, hops, *params):
        packed = model.packed_weights("fp32")
        dev = packed.device
        state = state.detach().to(device=dev, dtype=torch.float32).contiguous()
        x = x.detach().to(device=dev, dtype=torch.float32).contiguous()
        B, _, nx = state.shape
        L = model.num_layers
        with torch.cuda.device(dev):
            edges = torch.empty(B, 2 * hops * nx, dtype=torch.float32, device=dev)
            acts = torch.empty(_lib.lib().fluxgnn_train_acts_bytes(L, B, nx) // 4, dtype=torch.float32, device=dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(_lib.lib().fluxgnn_forward_ring_train(packed.data_ptr(), L, state.data_ptr(), x.data_ptr(), B, nx,
                                                             radius, hops, edges.data_ptr(), acts.data_ptr(), stream),
                       "fluxgnn_forward_ring_train")
        ctx.model, ctx.radius, ctx.hops = model, radius, hops
        # the parameters go through save_for_backward so that autograd's version check fires when one of
        # them is modified in place between this forward and the backward (optimizer.step() on a retained
        # graph, load_state_dict): stock autograd raises there, and so does this function
        ctx.save_for_backward(state, x, acts, *params)
        return edges

    @staticmethod
    def backward(ctx, dflux):
        model, radius, hops = ctx.model, ctx.radius, ctx.hops
        state, x, acts, *params = ctx.saved_tensors        # raises if a saved parameter changed in place
        dev = state.device
        B, _, nx = state.shape
        L, H = model.num_layers, model.hidden_dim
        f32 = lambda t: t.detach().to(torch.float32).contiguous()
        # model.parameters() order: input_mlp.0.{weight,bias}, update_mlps.l.0.{weight,bias}, edge_mlp.0.*, edge_mlp.2.*
        with torch.cuda.device(dev), torch.no_grad():
            w_in = f32(params[0])
            w_upd = torch.stack([f32(params[2 + 2 * l]) for l in range(L)]).contiguous()
            w_e1, w_e2 = f32(params[2 + 2 * L]), f32(params[4 + 2 * L])
            z = lambda *shape: torch.zeros(*shape, dtype=torch.float32, device=dev)
            g_w_in, g_b_in = z(H, model.input_dim), z(H)
            g_w_upd, g_b_upd = z(L, H, 2 * H), z(L, H)
            g_w_e1, g_b_e1, g_w_e2, g_b_e2 = z(H, 2 * H), z(H), z(1, H), z(1)
            dstate = torch.empty_like(state) if ctx.needs_input_grad[1] else None
            work = torch.empty(_lib.lib().fluxgnn_backward_workspace_bytes(B, nx) // 4, dtype=torch.float32, device=dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(_lib.lib().fluxgnn_backward_ring(
                w_in.data_ptr(), w_upd.data_ptr(), w_e1.data_ptr(), w_e2.data_ptr(), L,
                state.data_ptr(), x.data_ptr(), acts.data_ptr(), f32(dflux).data_ptr(), B, nx, radius, hops,
                g_w_in.data_ptr(), g_b_in.data_ptr(), g_w_upd.data_ptr(), g_b_upd.data_ptr(),
                g_w_e1.data_ptr(), g_b_e1.data_ptr(), g_w_e2.data_ptr(), g_b_e2.data_ptr(),
                dstate.data_ptr() if dstate is not None else None, work.data_ptr(), stream), "fluxgnn_backward_ring")
        # gradients in model.parameters() order: input_mlp, update_mlps[l], edge_mlp.0, edge_mlp.2
        grads = [g_w_in, g_b_in]
        for l in range(L):
            grads += [g_w_upd[l], g_b_upd[l]]
        grads += [g_w_e1, g_b_e1, g_w_e2, g_b_e2]
        return (None, dstate, None, None, None, *grads)


def ring_fluxes_with_grad(model, state, x, radius, hops):
    return RingFluxes.apply(model, state, x, radius, hops, *model.parameters())


def _param_views(params, L):
    """Weights the backward entry points read, from model.parameters() order:
    input_mlp.0.{weight,bias}, update_mlps.l.0.{weight,bias}, edge_mlp.0.*, edge_mlp.2.*"""
    f32 = lambda t: t.detach().to(torch.float32).contiguous()
    w_upd = torch.stack([f32(params[2 + 2 * l]) for l in range(L)]).contiguous()
    return f32(params[0]), w_upd, f32(params[2 + 2 * L]), f32(params[4 + 2 * L])


class HybridStep(torch.autograd.Function):
    """One differentiable hybrid step: (state'[B,3,nx], face_flux[B,nx]) from state[B,3,nx] -- the body of the
    reference's training rollout (scripts/training/train_ablation.py:172-206) as one fused forward launch
    (`fluxgnn_hybrid_step_train`) and one backward call (`fluxgnn_hybrid_step_backward`).  n' and u' are
    differentiable w.r.t. the state and the parameters; E' is the field solve of n' and carries NO gradient,
    as in the reference, where it goes through numpy (:198-200)."""

    @staticmethod
    def forward(ctx, model, grid, length, radius, c, dt, state, *params):
        packed = model.packed_weights("fp32")
        dev = packed.device
        state = state.detach().to(device=dev, dtype=torch.float32).contiguous()
        B, _, nx = state.shape
        L = model.num_layers
        x_dev, gtab = grid.tables(dev)
        lib = _lib.lib()
        with torch.cuda.device(dev):
            out = torch.empty_like(state)
            face = torch.empty(B, nx, dtype=torch.float32, device=dev)
            acts = torch.empty(lib.fluxgnn_train_acts_bytes(L, B, nx) // 4, dtype=torch.float32, device=dev)
            ws_bytes = lib.fluxgnn_hybrid_workspace_bytes(B, nx)
            work = torch.empty(ws_bytes // 4, dtype=torch.float32, device=dev) if ws_bytes else None
            _lib.check(lib.fluxgnn_hybrid_step_train(
                packed.data_ptr(), L, state.data_ptr(), out.data_ptr(), x_dev.data_ptr(),
                gtab.data_ptr() if gtab is not None else None, B, nx, length, radius, c, dt, face.data_ptr(),
                acts.data_ptr(), work.data_ptr() if work is not None else None,
                torch.cuda.current_stream(dev).cuda_stream), "fluxgnn_hybrid_step_train")
        ctx.model, ctx.radius, ctx.c, ctx.dt = model, radius, c, dt
        ctx.save_for_backward(state, x_dev, acts, *params)      # parameters: see RingFluxes.forward
        return out, face

    @staticmethod
    def backward(ctx, g_out, g_face):
        model = ctx.model
        state, x, acts, *params = ctx.saved_tensors
        dev = state.device
        B, _, nx = state.shape
        L, H = model.num_layers, model.hidden_dim
        lib = _lib.lib()
        with torch.cuda.device(dev), torch.no_grad():
            w_in, w_upd, w_e1, w_e2 = _param_views(params, L)
            z = lambda *shape: torch.zeros(*shape, dtype=torch.float32, device=dev)
            g_w_in, g_b_in = z(H, model.input_dim), z(H)
            g_w_upd, g_b_upd = z(L, H, 2 * H), z(L, H)
            g_w_e1, g_b_e1, g_w_e2, g_b_e2 = z(H, 2 * H), z(H), z(1, H), z(1)
            g_out = (z(B, 3, nx) if g_out is None else g_out.detach().to(torch.float32).contiguous())
            g_face = None if g_face is None else g_face.detach().to(torch.float32).contiguous()
            dstate = torch.empty_like(state)
            work = torch.empty(lib.fluxgnn_step_backward_workspace_bytes(B, nx) // 4, dtype=torch.float32, device=dev)
            _lib.check(lib.fluxgnn_hybrid_step_backward(
                w_in.data_ptr(), w_upd.data_ptr(), w_e1.data_ptr(), w_e2.data_ptr(), L, state.data_ptr(), x.data_ptr(),
                acts.data_ptr(), g_out.data_ptr(), g_face.data_ptr() if g_face is not None else None, B, nx, ctx.radius,
                ctx.c, ctx.dt, g_w_in.data_ptr(), g_b_in.data_ptr(), g_w_upd.data_ptr(), g_b_upd.data_ptr(),
                g_w_e1.data_ptr(), g_b_e1.data_ptr(), g_w_e2.data_ptr(), g_b_e2.data_ptr(), dstate.data_ptr(),
                work.data_ptr(), torch.cuda.current_stream(dev).cuda_stream), "fluxgnn_hybrid_step_backward")
        grads = [g_w_in, g_b_in]
        for l in range(L):
            grads += [g_w_upd[l], g_b_upd[l]]
        grads += [g_w_e1, g_b_e1, g_w_e2, g_b_e2]
        return (None, None, None, None, None, None, dstate if ctx.needs_input_grad[6] else None, *grads)


def hybrid_step_with_grad(model, grid, length, radius, c, dt, state):
    return HybridStep.apply(model, grid, length, radius, c, dt, state, *model.parameters())
