"""Phase timing of the tensor-core tile kernel (debug build: python -m gnn_plasma_flux_b200.build
-DFLUXGNN_TC_TIMING --suffix=_timing; run with FLUXGNN_LIB=.../libfluxgnn_timing.so)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gnn_plasma_flux_b200 import HybridSolver, _lib
from gnn_plasma_flux_b200.synthetic import seeded_model, stable_initial_conditions

prec = sys.argv[1] if len(sys.argv) > 1 else "fp16x3"
nx, B, r, steps = 64, 4096, 3, 20
dev = torch.device("cuda", 0)
sol = HybridSolver(None, r, nx=nx, dt=1e-3, device=dev, graph_radius=r, model=seeded_model(0, dev), precision=prec)
st = stable_initial_conditions(sol.baseline, B, distinct=64)
sol.rollout(st, 2)
torch.cuda.synchronize()
L = _lib.lib()
L.fluxgnn_debug_tc_timing.argtypes = [ctypes.c_void_p, ctypes.c_int]
buf = (ctypes.c_longlong * 16)()
L.fluxgnn_debug_tc_timing(None, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); sol.rollout(st, steps); e1.record(); torch.cuda.synchronize()
L.fluxgnn_debug_tc_timing(buf, 0)
# slots used by hybrid_tc16_kernel.cu (TC_TICK): 2 input layer (incl. next-tile load), 3 waiting for a group's UMMAs,
# 4 message-passing epilogues, 5 edge-readout epilogues, 8 finite-volume tail (face flux, update, field solve, write-out)
names = {2: "input layer", 3: "wait UMMA", 4: "layer epilogues", 5: "edge readout", 8: "finite-volume tail"}
tiles = (B * nx // 256 + 147) // 148     # assumes a 148-CTA grid
tot = sum(buf[i] for i in names)
print(f"{prec}: {e0.elapsed_time(e1) / steps:.3f} ms/step; CTA 0: {tiles} tiles x {steps} steps, {tot / (tiles * steps):.0f} clk per tile-step")
for i, n in names.items():
    print(f"  {n:16s} {buf[i] / (tiles * steps):9.0f} clk/tile-step  {100 * buf[i] / tot:5.1f} %")
