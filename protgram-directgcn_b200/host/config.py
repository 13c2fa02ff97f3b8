"""Config attribute bag -- the hot-path subset of the reference's config.py:13-172 (same
attribute names and defaults; callers mutate attributes after construction, as in
run_graph_builder.py:36-43 and unit_tests.py:63-68)."""
from __future__ import annotations

import os
from pathlib import Path
from typing import Dict, Optional


class Config:
    def __init__(self):
        self.RANDOM_STATE = 42
        self.DEBUG_VERBOSE = False
        self.PROJECT_ROOT = Path(".").resolve()
        self.BASE_DATA_DIR = self.PROJECT_ROOT / "data"
        self.BASE_OUTPUT_DIR = self.BASE_DATA_DIR / "pipeline_output"
        self.GCN_INPUT_FASTA_PATH = self.BASE_DATA_DIR / "sequences.fasta"
        self.GRAPH_OBJECTS_DIR = self.BASE_OUTPUT_DIR / "1_graph_objects"
        # --- GCN pipeline parameters (reference config.py:60-104) ---
        self.GCN_NGRAM_MAX_N = 3
        self.GRAPH_BUILDER_WORKERS: Optional[int] = max(1, (os.cpu_count() or 5) - 4)
        self.GCN_HIDDEN_LAYER_DIMS = [256, 128, 64]
        self.GCN_1GRAM_INIT_DIM = 512
        self.GCN_EPOCHS_PER_LEVEL = 500
        self.GCN_LR = 0.001
        self.GCN_DROPOUT_RATE = 0.5
        self.GCN_WEIGHT_DECAY = 1e-4
        self.GCN_L2_REG_LAMBDA = 1e-7
        self.GCN_PROPAGATION_EPSILON = 1e-9
        self.GCN_MAX_PE_LEN = 512
        self.GCN_USE_VECTOR_COEFFS = True
        self.GCN_TASK_TYPES_PER_LEVEL: Dict[int, str] = {1: "next_node", 2: "next_node", 3: "next_node"}
        self.GCN_DEFAULT_TASK_TYPE = "community"
        self.GCN_USE_CLUSTER_TRAINING = True
        self.GCN_CLUSTER_TRAINING_THRESHOLD_NODES = 10000
