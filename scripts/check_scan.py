"""Scan field solve vs the FFT solve: deviation, certificate verdicts and step time (one B200)."""
import sys, time
import numpy as np
import torch
from gnn_plasma_flux_b200 import BaselineSolver, _lib
from gnn_plasma_flux_b200.synthetic import stable_initial_conditions


def rel(a, b):
    return [float((a[:, c] - b[:, c]).abs().max() / b[:, c].abs().max()) for c in range(3)]


def main():
    for nx, B, steps in [(4096, 2, 20), (12000, 1, 20), (1 << 16, 3, 20), (1 << 20, 2, 20), (1 << 24, 1, 20), (1 << 22, 8, 5)]:
        dx = 2 * np.pi / nx
        dt = min(0.02 * dx, 0.2 * dx * dx / 1e-3)       # advective and viscous stability
        sol = BaselineSolver(nx=nx, dt=dt, nu=1e-3, device="cuda")
        state = stable_initial_conditions(sol, B)
        ref = sol.rollout(state, steps, field_solve="spectral")[0]
        got = sol.rollout(state, steps, field_solve="auto")[0]
        print(f"nx={nx} B={B} steps={steps}: ran {sol.last_field_solve}, first uncertified {sol.last_uncertified_step}, "
              f"rel dev n,u,E = {rel(got, ref)}", flush=True)
        one_s = sol.rollout(state, 1, field_solve="spectral")[0]
        one = sol.rollout(state, 1, field_solve="auto")[0]
        print("   one step: n,u bit-identical:", bool(torch.equal(one[:, :2], one_s[:, :2])), "E rel", rel(one, one_s)[2])
        for mode in ("spectral", "auto"):
            for _ in range(2):
                sol.rollout(state, steps, field_solve=mode)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                sol.rollout(state, steps, field_solve=mode)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3 / steps
            print(f"   {mode}: {ms*1e3:.1f} us/step, {B*nx/ms/1e6:.2f} Gcell-updates/s, "
                  f"{B*nx*24/ms/1e6/6543.1*100:.1f} % of the 24 B/cell roof")
    # rough density: certificate must fail, auto must fall back bit-identically
    nx = 1 << 16
    sol = BaselineSolver(nx=nx, dt=0.02 * (2 * np.pi / nx), nu=1e-3, device="cuda")
    state = stable_initial_conditions(sol, 2)
    state[:, 0] += 1e-3 * torch.randn(2, nx, device="cuda")
    ref = sol.rollout(state, 5, field_solve="spectral")[0]
    got = sol.rollout(state, 5, field_solve="auto")[0]
    print("white-noise density: ran", sol.last_field_solve, "first uncertified", sol.last_uncertified_step,
          "identical to spectral:", bool(torch.equal(got, ref)))
    try:
        sol.rollout(state, 5, field_solve="scan")
        print("scan did not raise (unexpected)")
    except _lib.FluxGNNError as e:
        print("scan raised:", str(e)[:100])
    # trajectory + flux
    nx = 1 << 14
    sol = BaselineSolver(nx=nx, dt=0.02 * (2 * np.pi / nx), nu=1e-3, device="cuda")
    state = stable_initial_conditions(sol, 2)
    for rec in (1, 3):
        o1, t1, f1 = sol.rollout(state, 9, record_every=rec, record_flux=True, field_solve="spectral")
        o2, t2, f2 = sol.rollout(state, 9, record_every=rec, record_flux=True, field_solve="scan")
        print(f"record_every={rec}: traj dev", float((t1 - t2).abs().max()), "flux dev", float((f1 - f2).abs().max()),
              "final dev", float((o1 - o2).abs().max()), "traj[-1]==final", bool(torch.equal(t2[-1], o2)))


if __name__ == "__main__":
    main()
