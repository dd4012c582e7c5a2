"""ctypes binding of libfluxgnn.so (C ABI in include/fluxgnn.h).

There is no CPU or PyTorch fallback: if the shared library is missing or a call
fails, the error is raised to the caller.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_longlong, c_size_t, c_ulonglong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FLUXGNN_LIB") or os.path.join(_HERE, "libfluxgnn.so")   # FLUXGNN_LIB: experiment builds

ABI_VERSION = 4
MAX_HOPS = 4
MAX_LAYERS = 8
HIDDEN = 128
INPUT_DIM = 4
GENERIC_HIDDEN = (16, 32, 64, 128)     # other architectures: the generic FP32-pipe kernels (csrc/generic_kernels.cu)
GENERIC_MAX_INPUT_DIM = 16
# fp16 / fp16x3 operand images store 2^8 * W (kTc16WeightScale, csrc/common.cuh): fp16's largest finite value / 2^8
FP16_WEIGHT_LIMIT = 65504.0 / 256.0

TC_PRECISIONS = {"tf32x3": 1, "tf32": 2, "fp16x3": 3, "fp16": 4, "bf16": 5}     # FLUXGNN_TC_* of include/fluxgnn.h


def weight_layout(precision: str) -> str:
    """Packed-weight layout a precision mode streams: the FP32-pipe kernel's K-major halves, the
    TF32 UMMA operand images, or the 16-bit images (fp16 serves fp16x3 and fp16; bf16 its own)."""
    return {"fp32": "fp32", "tf32x3": "tc", "tf32": "tc", "fp16x3": "tc16", "fp16": "tc16", "bf16": "tc16_bf16"}[precision]

# name -> (restype, argtypes); mirrors include/fluxgnn.h declaration by declaration
SIGNATURES = {
    "fluxgnn_abi_version": (c_int, []),
    "fluxgnn_last_error": (c_char_p, []),
    "fluxgnn_launch_count": (c_ulonglong, []),
    "fluxgnn_latency_cluster_slots": (c_int, []),
    "fluxgnn_ffma_probe": (c_longlong, [c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "fluxgnn_packed_weight_bytes": (c_size_t, [c_int]),
    "fluxgnn_pack_weights": (c_int, [c_void_p] * 8 + [c_int, c_void_p, c_void_p]),
    "fluxgnn_poisson_table": (c_int, [c_int, c_double, c_void_p, c_void_p]),
    "fluxgnn_poisson_uses_table": (c_int, [c_int]),
    "fluxgnn_poisson_workspace_bytes": (c_size_t, [c_int, c_int]),
    "fluxgnn_poisson_spectral": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_void_p, c_int, c_int, c_double,
                                         c_void_p, c_void_p]),
    "fluxgnn_forward_ring": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                     c_void_p, c_void_p, c_void_p]),
    "fluxgnn_hybrid_workspace_bytes": (c_size_t, [c_int, c_int]),
    "fluxgnn_hybrid_rollout": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_double,
                                       c_int, c_float, c_float, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "fluxgnn_hybrid_rollout_diag": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                            c_double, c_int, c_float, c_float, c_int, c_void_p, c_void_p]),
    "fluxgnn_packed_tc_weight_bytes": (c_size_t, [c_int]),
    "fluxgnn_pack_weights_tc": (c_int, [c_void_p] * 8 + [c_int, c_void_p, c_void_p]),
    "fluxgnn_packed_tc16_weight_bytes": (c_size_t, [c_int]),
    "fluxgnn_pack_weights_tc16": (c_int, [c_void_p] * 8 + [c_int, c_int, c_void_p, c_void_p]),
    "fluxgnn_forward_ring_tc": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int,
                                        c_void_p, c_void_p, c_void_p]),
    "fluxgnn_hybrid_rollout_tc": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                          c_double, c_int, c_float, c_float, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "fluxgnn_train_acts_bytes": (c_size_t, [c_int, c_int, c_int]),
    "fluxgnn_forward_ring_train": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                           c_void_p, c_void_p, c_void_p]),
    "fluxgnn_backward_workspace_bytes": (c_size_t, [c_int, c_int]),
    "fluxgnn_backward_ring": (c_int, [c_void_p] * 4 + [c_int] + [c_void_p] * 4 + [c_int] * 4 + [c_void_p] * 11),
    "fluxgnn_hybrid_step_train": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_double,
                                          c_int, c_float, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fluxgnn_step_backward_workspace_bytes": (c_size_t, [c_int, c_int]),
    "fluxgnn_hybrid_step_backward": (c_int, [c_void_p] * 4 + [c_int] + [c_void_p] * 5 + [c_int] * 3 + [c_float] * 2
                                     + [c_void_p] * 11),
    "fluxgnn_rollout_metrics": (c_int, [c_void_p, c_void_p, c_longlong, c_int, c_void_p, c_void_p]),
    "fluxgnn_hybrid_slab_step": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                         c_float, c_float, c_void_p]),
    "fluxgnn_hybrid_slab_step_ld": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                            c_int, c_int, c_float, c_float, c_void_p]),
    "fluxgnn_hybrid_slab_step_peer": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                              c_int, c_int, c_float, c_float, c_void_p, c_void_p, c_void_p]),
    "fluxgnn_baseline_slab_step_peer": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int,
                                                c_float, c_float, c_float, c_float, c_void_p, c_void_p, c_void_p]),
    "fluxgnn_scan_slab_field_peer": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_int, c_int, c_int, c_int, c_double,
                                             c_void_p, c_void_p, c_double, c_int, c_void_p, c_void_p, c_void_p, c_int,
                                             c_void_p]),
    "fluxgnn_scan_slab_sums_peer": (c_int, [c_void_p, c_longlong, c_int, c_int, c_longlong, c_void_p, c_void_p, c_void_p,
                                            c_longlong, c_int, c_int, c_void_p]),
    "fluxgnn_peer_halo_push": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "fluxgnn_peer_allgather": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_int, c_int, c_void_p]),
    "fluxgnn_baseline_slab_step": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int,
                                           c_float, c_float, c_float, c_float, c_void_p]),
    "fluxgnn_poisson_dist_pack": (c_int, [c_void_p, c_longlong, c_int, c_int, c_void_p, c_void_p]),
    "fluxgnn_poisson_dist_unpack": (c_int, [c_void_p, c_void_p, c_longlong, c_int, c_int, c_void_p]),
    "fluxgnn_poisson_dist_rank_dft": (c_int, [c_void_p, c_void_p, c_int, c_longlong, c_longlong, c_int, c_int, c_void_p]),
    "fluxgnn_poisson_dist_local": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_double, c_void_p]),
    "fluxgnn_generic_packed_bytes": (c_size_t, [c_int, c_int, c_int]),
    "fluxgnn_generic_pack": (c_int, [c_void_p] * 8 + [c_int, c_int, c_int, c_void_p, c_void_p]),
    "fluxgnn_generic_forward_ring": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                             c_int, c_void_p, c_void_p, c_void_p]),
    "fluxgnn_generic_workspace_bytes": (c_size_t, [c_int, c_int]),
    "fluxgnn_generic_hybrid_rollout": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                               c_double, c_int, c_float, c_float, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "fluxgnn_pure_gnn_packed_bytes": (c_size_t, [c_int, c_int]),
    "fluxgnn_pure_gnn_pack": (c_int, [c_void_p] * 8 + [c_int, c_int, c_void_p, c_void_p]),
    "fluxgnn_pure_gnn_rollout": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "fluxgnn_pure_gnn_delta": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "fluxgnn_dense_layer": (c_int, [c_void_p] * 5 + [c_int] * 4 + [c_void_p]),
    "fluxgnn_baseline_scan_supported": (c_int, [c_int, c_int]),
    "fluxgnn_baseline_scan_workspace_bytes": (c_size_t, [c_int, c_int]),
    "fluxgnn_baseline_rollout_scan": (c_int, [c_void_p, c_void_p, c_int, c_int, c_double, c_float, c_float, c_float, c_float,
                                              c_int, c_int, c_void_p, c_void_p, c_double, c_void_p, c_void_p, c_void_p]),
    "fluxgnn_scan_slab_supported": (c_int, [c_int, c_int]),
    "fluxgnn_scan_slab_workspace_bytes": (c_size_t, [c_int, c_int]),
    "fluxgnn_scan_slab_sums": (c_int, [c_void_p, c_longlong, c_int, c_int, c_longlong, c_void_p, c_void_p, c_void_p]),
    "fluxgnn_scan_slab_field": (c_int, [c_void_p, c_longlong, c_void_p, c_longlong, c_int, c_int, c_int, c_int, c_double,
                                        c_void_p, c_void_p, c_double, c_int, c_void_p, c_void_p]),
    "fluxgnn_scan_slab_certify": (c_int, [c_int, c_int, c_int, c_double, c_void_p, c_double, c_int, c_void_p, c_void_p]),
    "fluxgnn_baseline_workspace_bytes": (c_size_t, [c_int, c_int]),
    "fluxgnn_baseline_rollout": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_double, c_float, c_float, c_float,
                                         c_float, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
}


class FluxGNNError(RuntimeError):
    """A libfluxgnn entry point returned a negative status."""


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FluxGNNError(
                f"{LIB_PATH} not found: build it with `python -m gnn_plasma_flux_b200.build` "
                "(nvcc, sm_100a). There is no CPU fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        handle.fluxgnn_abi_version.restype = c_int
        if handle.fluxgnn_abi_version() != ABI_VERSION:          # checked first: a stale library lacks newer symbols
            raise FluxGNNError(f"{LIB_PATH} has ABI version {handle.fluxgnn_abi_version()}, this package needs "
                               f"{ABI_VERSION}: rebuild it with `python -m gnn_plasma_flux_b200.build`")
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().fluxgnn_last_error().decode("utf-8", "replace")
        raise FluxGNNError(f"{what} failed ({status}): {msg}")


def launch_count() -> int:
    return int(lib().fluxgnn_launch_count())
