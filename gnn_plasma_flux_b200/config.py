"""Constants of the reference's ablation study (data only; src/config.py:9-79).

Only the entries the hot path reads are reproduced: HybridSolver always builds
its model from MODEL_CONFIG (src/hybrid_solver.py:21-26).
"""

DATASET_CONFIG = {"nx": 64, "num_initial_conditions": 50, "steps_per_ic": 40, "dt": 5e-3, "t_end": 1.0, "nu": 1e-3}
MODEL_CONFIG = {"input_dim": 4, "hidden_dim": 128, "num_layers": 4}
STENCIL_RADII = [1, 2, 3]
EVAL_CONFIG = {"n_steps": 100, "test_seed": 123}
