"""Build libfluxgnn.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension).

    python -m gnn_plasma_flux_b200.build [--force]

The shared library lands next to this file so that it travels with the source
tree; it is git-ignored.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libfluxgnn.so")
SOURCES = ["api.cu", "field_kernels.cu", "fft_poisson.cu", "scan_poisson.cu", "hybrid_kernel.cu", "hybrid_tc_kernel.cu", "hybrid_tc16_kernel.cu", "train_kernels.cu", "comparison_kernels.cu", "generic_kernels.cu", "hybrid_latency_kernel.cu"] + [f"hybrid_r{r}.cu" for r in range(5)] + [f"hybrid_train_r{r}.cu" for r in range(5)] + [f"hybrid_cluster_r{r}.cu" for r in range(1, 5)]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]


def _nvcc() -> str:
    cand = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")
    return cand if os.path.exists(cand) else "nvcc"


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False, defines: tuple[str, ...] = (), suffix: str = "") -> str:
    """`defines`/`suffix` build an experiment variant (e.g. -DFLUXGNN_FFT_COL_CTAS=3 -> libfluxgnn_x.so,
    loaded with FLUXGNN_LIB=...); the default call builds the product library."""
    OBJ = os.path.join(HERE, "build" + suffix)
    LIB = os.path.join(HERE, f"libfluxgnn{suffix}.so")
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "fluxgnn.h"))

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        path = os.path.join(CSRC, src)
        if force or _stale(obj, [path] + headers):
            cmd = [_nvcc(), *NVCC_FLAGS, *defines, "-c", path, "-o", obj]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
            if res.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}")
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    if force or _stale(LIB, objs):
        cmd = [_nvcc(), "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("link of libfluxgnn.so failed")
    return LIB


if __name__ == "__main__":
    defs = tuple(a for a in sys.argv[1:] if a.startswith("-D"))
    sfx = next((a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--suffix=")), "")
    print(build_library(force="--force" in sys.argv, verbose="--quiet" not in sys.argv, defines=defs, suffix=sfx))
