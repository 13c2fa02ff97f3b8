"""protgram-directgcn_b200: B200 (sm_100a) implementation of ProtGram-DirectGCN's two hot paths.

    host/   same-named mirror of the reference's Python API for the hot path
            (GraphBuilder, DirectedNgramGraph, DirectGCNLayer, ProtGramDirectGCN, ...)
    csrc/   hand-written CUDA kernels behind the C ABI of include/pgb200.h (libpgb200.so)

Import name: `protgram_directgcn_b200` (see protgram_directgcn_b200.py at the repo root).
"""
from . import _native  # noqa: F401
from .host.config import Config  # noqa: F401
from .host.data_utils import DataLoader, DataUtils  # noqa: F401
from .host.graph_utils import DirectedNgramGraph, Graph  # noqa: F401
from .host.data_builder import GraphBuilder  # noqa: F401
from .host.protgram_directgcn import Data, DirectGCNLayer, ProtGramDirectGCN  # noqa: F401
from .host.models_utils import EmbeddingProcessor  # noqa: F401
from .host.trainer_utils import create_clustered_subgraphs, generate_next_node_labels, init_level_features  # noqa: F401

__all__ = ["Config", "DataLoader", "DataUtils", "Graph", "DirectedNgramGraph", "GraphBuilder", "Data",
           "DirectGCNLayer", "ProtGramDirectGCN", "EmbeddingProcessor", "generate_next_node_labels", "init_level_features",
           "create_clustered_subgraphs"]
