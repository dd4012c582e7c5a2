"""Rollout diagnostics on the device.

The reference's evaluation scripts copy every trajectory to the host and reduce it with
numpy (scripts/evaluation/evaluate_all.py:118-159: per-step MSE per channel, energy
0.5*mean(u^2+E^2) and its drift, charge mean(n) and its drift; evaluate_long_rollout.py:53-66:
first non-finite step).  Here one kernel reduces device-resident trajectories, so only
[T, B, 8] floats ever cross PCIe.
"""
from __future__ import annotations

import torch

from . import _lib


def rollout_metrics(pred: torch.Tensor, truth: torch.Tensor | None = None) -> dict:
    """pred, truth: [..., 3, nx] CUDA float32 (e.g. [T,B,3,nx]).  Returns tensors shaped like the
    leading axes: mse_n, mse_u, mse_E, mse_total, energy, charge, nonfinite (count)."""
    if pred.shape[-2] != 3:
        raise ValueError(f"states must be [...,3,nx], got {tuple(pred.shape)}")
    if not pred.is_cuda:
        raise _lib.FluxGNNError("rollout_metrics needs CUDA tensors (there is no CPU fallback)")
    pred = pred.to(torch.float32).contiguous()
    if truth is not None:
        if truth.shape != pred.shape:
            raise ValueError("pred and truth must have the same shape")
        truth = truth.to(device=pred.device, dtype=torch.float32).contiguous()
    lead, nx = pred.shape[:-2], pred.shape[-1]
    n_states = 1
    for d in lead:
        n_states *= int(d)
    with torch.cuda.device(pred.device):
        out = torch.empty(n_states, 8, dtype=torch.float32, device=pred.device)
        stream = torch.cuda.current_stream(pred.device).cuda_stream
        _lib.check(_lib.lib().fluxgnn_rollout_metrics(pred.data_ptr(), truth.data_ptr() if truth is not None else None,
                                                      n_states, nx, out.data_ptr(), stream), "fluxgnn_rollout_metrics")
    out = out.reshape(*lead, 8)
    res = {"mse_n": out[..., 0], "mse_u": out[..., 1], "mse_E": out[..., 2], "energy": out[..., 3],
           "charge": out[..., 4], "nonfinite": out[..., 5]}
    res["mse_total"] = res["mse_n"] + res["mse_u"] + res["mse_E"]
    return res


def compute_metrics(states_pred: torch.Tensor, states_true: torch.Tensor) -> dict:
    """Device version of scripts/evaluation/evaluate_all.py:118-159 for trajectories [T,3,nx]
    (or batched [T,B,3,nx]): same keys; time series stay on the device as tensors."""
    p, t = rollout_metrics(states_pred, states_true), rollout_metrics(states_true)
    e_drift_pred = (p["energy"] - p["energy"][0]).abs()
    e_drift_true = (t["energy"] - t["energy"][0]).abs()
    c_drift_pred = (p["charge"] - p["charge"][0]).abs()
    c_drift_true = (t["charge"] - t["charge"][0]).abs()
    return {"mse_n": p["mse_n"], "mse_u": p["mse_u"], "mse_E": p["mse_E"], "mse_total": p["mse_total"],
            "energy_drift_pred": e_drift_pred, "energy_drift_true": e_drift_true,
            "charge_drift_pred": c_drift_pred, "charge_drift_true": c_drift_true,
            "final_mse": p["mse_total"][-1], "mean_mse": p["mse_total"].mean(dim=0),
            "final_energy_drift": e_drift_pred[-1], "final_charge_drift": c_drift_pred[-1]}


def first_nonfinite_step(traj: torch.Tensor) -> torch.Tensor:
    """For a trajectory [T,B,3,nx]: per IC the first stored step with a NaN/Inf, or -1
    (the `exploded_at` bookkeeping of evaluate_long_rollout.py:53-66)."""
    bad = rollout_metrics(traj)["nonfinite"] > 0                       # [T,B]
    T = bad.shape[0]
    idx = torch.where(bad, torch.arange(T, device=bad.device).unsqueeze(-1).expand_as(bad),
                      torch.full_like(bad, T, dtype=torch.long))
    first = idx.min(dim=0).values
    return torch.where(first == T, torch.full_like(first, -1), first)
