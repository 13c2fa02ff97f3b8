"""Shared helpers for the parity tests (fixtures loading, COO comparison)."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
BUILD_FIXTURES = {"build_ka1": 3, "build_ka2": 1, "build_ragged": 4, "build_protein": 3,
                  "build_lowcomplexity": 3}
MODEL_FIXTURES = ["model_refgraph", "model_n1_pe", "model_general_scalar", "model_cluster_batch"]
MATS = ("A_out_w", "A_in_w", "A_undirected_norm_sparse", "mathcal_A_out", "mathcal_A_in")


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def fasta_sequences(fasta_text, tmp_path):
    """Write the fixture's FASTA text and return its path."""
    p = os.path.join(str(tmp_path), "in.fasta")
    with open(p, "w") as f:
        f.write(fasta_text)
    return p


def golden_edges(g, n):
    idx = g[f"n{n}_A_out_w_idx"]
    return idx[0], idx[1], g[f"n{n}_A_out_w_val"]


def rel_err(a, b):
    """norm-wise relative error |a-b|_inf / max(|b|_inf, tiny)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0 and b.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-30))
