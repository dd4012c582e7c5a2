import ctypes, os, sys
sys.path.insert(0, "/root/repo")
import torch
from gnn_plasma_flux_b200 import _lib
L = _lib.lib()
torch.zeros(1, device="cuda")
L.fluxgnn_debug_max_clusters.restype = ctypes.c_int
L.fluxgnn_debug_max_clusters.argtypes = [ctypes.c_int]
print({c: L.fluxgnn_debug_max_clusters(c) for c in (2, 3, 4, 5, 6, 7, 8)})
