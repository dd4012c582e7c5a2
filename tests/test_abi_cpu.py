"""CPU suite: the C-ABI library builds, loads, exports every symbol include/fluxgnn.h
declares, and the host-side drop-in classes keep the reference's surface.  No
compute entry point is called here (no GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT


def header_functions():
    text = open(os.path.join(ROOT, "include", "fluxgnn.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fluxgnn_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_exported(built_lib):
    from gnn_plasma_flux_b200 import _lib
    names = header_functions()
    assert len(names) >= 10
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in names:
        assert hasattr(raw, name), f"{name} declared in fluxgnn.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes SIGNATURES out of sync with fluxgnn.h"
    assert built_lib.fluxgnn_abi_version() == _lib.ABI_VERSION


def test_precision_codes_match_header(built_lib):
    """FLUXGNN_TC_* in the header == the precision names the Python classes accept; every mode has a
    packed-weight layout and a size query that answers without a GPU."""
    from gnn_plasma_flux_b200 import _lib
    text = open(os.path.join(ROOT, "include", "fluxgnn.h")).read()
    codes = {name.lower(): int(val) for name, val in re.findall(r"#define\s+FLUXGNN_TC_([A-Z0-9]+)\s+(\d+)", text)}
    assert codes == _lib.TC_PRECISIONS
    for name in ["fp32", *codes]:
        assert _lib.weight_layout(name) in ("fp32", "tc", "tc16", "tc16_bf16")
    small = 2048 * 4
    assert built_lib.fluxgnn_packed_tc16_weight_bytes(4) == small + 5 * 8 * 16384           # 8 units of 16 KiB per layer
    assert built_lib.fluxgnn_packed_tc_weight_bytes(4) == small + 5 * 16 * 16384
    assert built_lib.fluxgnn_packed_tc16_weight_bytes(0) == 0
    rc = built_lib.fluxgnn_pack_weights_tc16(None, None, None, None, None, None, None, None, 4, 2, None, None)
    assert rc == -1                                               # precision 2 (tf32) has no 16-bit image
    assert built_lib.fluxgnn_pure_gnn_packed_bytes(64, 3) > 0 and built_lib.fluxgnn_pure_gnn_packed_bytes(48, 3) == 0


def test_size_queries(built_lib):
    small, layer = 2048, 2 * 128 * 128
    for L in (1, 4, 8):
        assert built_lib.fluxgnn_packed_weight_bytes(L) == 4 * (small + (L + 1) * layer)
    assert built_lib.fluxgnn_packed_weight_bytes(0) == 0
    assert built_lib.fluxgnn_packed_weight_bytes(9) == 0
    assert built_lib.fluxgnn_hybrid_workspace_bytes(16, 64) == 0
    assert built_lib.fluxgnn_hybrid_workspace_bytes(16, 1024) == 16 * 3 * 1024 * 4        # FFT inside one CTA: no scratch
    assert built_lib.fluxgnn_poisson_workspace_bytes(2, 1 << 15) == 0                      # nx/2 complex points fit one CTA
    assert built_lib.fluxgnn_poisson_workspace_bytes(2, 1 << 16) == 2 * (1 << 16) * 4      # four-step FFT: nx/2 complex64 of scratch
    assert built_lib.fluxgnn_hybrid_workspace_bytes(2, 1 << 16) == 2 * (1 << 16) * (12 + 4)
    assert built_lib.fluxgnn_baseline_workspace_bytes(2, 1000) == 2 * 3 * 1000 * 4
    assert [built_lib.fluxgnn_poisson_uses_table(n) for n in (64, 128, 256, 1000, 1024, 1 << 20)] == [1, 1, 0, 1, 0, 0]


def test_argument_errors_without_gpu(built_lib):
    from gnn_plasma_flux_b200 import _lib
    rc = built_lib.fluxgnn_pack_weights(None, None, None, None, None, None, None, None, 4, None, None)
    assert rc == -1 and b"null" in built_lib.fluxgnn_last_error()
    rc = built_lib.fluxgnn_hybrid_rollout(None, 4, None, None, None, None, 1, 64, 6.28, 1, 0.1, 0.1, 1, 1, None, None, None)
    assert rc == -1
    rc = built_lib.fluxgnn_forward_ring(ctypes.c_void_p(16), 12, None, None, 1, 64, 1, 1, None, None, None)
    assert rc == -3                                               # unsupported layer count
    with pytest.raises(_lib.FluxGNNError):
        _lib.check(rc, "forward_ring")


def test_fluxgnn_module_surface(weights):
    from gnn_plasma_flux_b200 import FluxGNN, MODEL_CONFIG
    torch.manual_seed(0)
    m = FluxGNN(**MODEL_CONFIG)
    sd = m.state_dict()
    assert set(sd) == set(weights)
    for key, val in sd.items():                                    # same construction order => same seeded init
        np.testing.assert_array_equal(val.numpy(), weights[key])
    m.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()})
    d = FluxGNN()
    assert (d.input_dim, d.hidden_dim, d.num_layers) == (2, 32, 2)     # reference defaults, src/flux_gnn.py:11
    from gnn_plasma_flux_b200 import _lib
    with pytest.raises(_lib.FluxGNNError):                         # loud failure, no CPU fallback
        m(torch.zeros(64, 4), torch.from_numpy(__import__("oracle.ref_port", fromlist=["x"]).ring_edges(64, 1)))
    with pytest.raises(NotImplementedError):                       # arbitrary graphs: documented deviation
        m(torch.zeros(64, 4), torch.zeros(2, 128, dtype=torch.long))


def test_build_chain_graph_matches_oracle():
    from gnn_plasma_flux_b200 import build_chain_graph
    from gnn_plasma_flux_b200.graph_constructor import is_ring
    from oracle import ref_port as P
    rng = np.random.RandomState(0)
    state = rng.randn(3, 64)                                      # float64 in, as examples/smoke_test.py:36
    x = np.linspace(0, 1, 64)
    nf, ei = build_chain_graph(state, x, "cpu")
    assert nf.shape == (64, 4) and nf.dtype == torch.float32
    assert ei.shape == (2, 128) and ei.dtype == torch.long
    np.testing.assert_array_equal(nf.numpy(), P.node_features(state, x))
    np.testing.assert_array_equal(ei.numpy(), P.ring_edges(64, 1))
    for r in (2, 3):
        _, e = build_chain_graph(torch.from_numpy(state.astype(np.float32)), x, radius=r)
        np.testing.assert_array_equal(e.numpy(), P.ring_edges(64, r))
        assert is_ring(e.clone(), 64) == r                         # untagged copy is recognised by value
    assert is_ring(torch.randint(0, 64, (2, 128)), 64) is None


def test_baseline_host_surface():
    from gnn_plasma_flux_b200.grid import PeriodicGrid
    from oracle import ref_port as P
    for nx in (64, 1000):
        g, o = PeriodicGrid(nx), P.Grid(nx=nx)
        np.testing.assert_array_equal(g.x, o.x)
        np.testing.assert_array_equal(g.k, o.k)
        assert g.dx == o.dx


def test_emulated_fabric_blocks_alias_and_offsets():
    """Host logic of the peer-memory fabric (domain.EmulatedFabric, the one-process stand-in for symmetric memory): a
    view of rank r's block aliases that rank's allocation at the given byte offset, whichever rank asks for it."""
    import torch
    from gnn_plasma_flux_b200.domain import EmulatedFabric
    fabrics = EmulatedFabric.create(3, "cpu")
    blocks = [f.allocate(4096) for f in fabrics]
    assert all(b.local.numel() == 4096 and b.local.dtype == torch.uint8 for b in blocks)
    mine = blocks[1].view(1, 256, (2, 3, 8))                    # rank 1's own view
    theirs = blocks[0].view(1, 256, (2, 3, 8))                  # rank 0 looking at rank 1's block
    assert mine.dtype == torch.float32 and mine.data_ptr() == theirs.data_ptr() == blocks[1].local.data_ptr() + 256
    theirs.fill_(7.0)
    assert float(mine.sum()) == 7.0 * 48
    assert float(blocks[2].view(2, 256, (48,)).abs().sum()) == 0.0     # another rank's block is untouched
    msg = blocks[2].view(0, 1024, (3, 96), torch.uint8)
    assert msg.data_ptr() == blocks[0].local.data_ptr() + 1024 and msg.shape == (3, 96)
    second = [f.allocate(64) for f in fabrics]                  # allocations pair up by order
    assert second[0].view(2, 0, (16,)).data_ptr() == second[2].local.data_ptr()
