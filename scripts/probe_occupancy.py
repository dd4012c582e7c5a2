"""FP32-pipe probe at different warps-per-scheduler counts (how well few warps feed the FMA pipe)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_plasma_flux_b200 import _lib
dev = torch.device("cuda")
sink = torch.empty(148 * 8 * 256, device=dev)
st = torch.cuda.current_stream()
for blocks, threads in [(148, 128), (148, 256), (296, 256), (148 * 8, 256)]:
    for packed in (0, 1):
        best = 0
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fl = _lib.lib().fluxgnn_ffma_probe(sink.data_ptr(), blocks, threads, 40000, packed, st.cuda_stream)
            e1.record(); torch.cuda.synchronize()
            best = max(best, fl / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        print(f"blocks/SM={blocks/148:.0f} threads={threads} warps/SMSP={blocks/148*threads/128:.0f} {'FFMA2' if packed else 'FFMA '} {best:.1f} TFLOP/s")
