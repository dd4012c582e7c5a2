"""gnn_plasma_flux_b200 -- the hybrid rollout hot path of gnn-plasma-flux on B200 (sm_100a).

Drop-in replacements for the reference's `src.flux_gnn.FluxGNN`,
`src.graph_constructor.build_chain_graph`, `src.hybrid_solver.HybridSolver` and
`src.baseline_solver.BaselineSolver`, backed by hand-written CUDA in
libfluxgnn.so (C ABI: include/fluxgnn.h).  Build the library first:

    python -m gnn_plasma_flux_b200.build
"""
from .baseline_solver import BaselineSolver
from .comparison_models import PINN, PureGNN
from .config import DATASET_CONFIG, EVAL_CONFIG, MODEL_CONFIG, STENCIL_RADII
from .flux_gnn import FluxGNN
from .graph_constructor import build_chain_graph, ring_edge_index
from .hybrid_solver import HybridSolver
from .datagen import generate_dataset
from .metrics import compute_metrics, first_nonfinite_step, rollout_metrics

__all__ = ["BaselineSolver", "FluxGNN", "HybridSolver", "build_chain_graph", "ring_edge_index",
           "generate_dataset", "PureGNN", "PINN", "compute_metrics", "rollout_metrics", "first_nonfinite_step",
           "DATASET_CONFIG", "EVAL_CONFIG", "MODEL_CONFIG", "STENCIL_RADII"]
