"""CPU oracle for the rows SURVEY.md section 8 marks "next" (f2, f4) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may import this package; the
product path never does.

Fresh restatement of the reference algorithm (citations into /root/reference):

  pool_proteins         <- src/utils/models_utils.py:210-262 (EmbeddingProcessor.pool_ngram_embeddings_for_protein_fast)
  init_level_features   <- src/pipeline/protgram_directgcn_trainer.py:312-330 (features of level n from level n-1)
  next_node_label_sets  <- src/pipeline/protgram_directgcn_trainer.py:222-237 (_generate_next_node_labels)

Parity status: pool_proteins and next_node_label_sets are PINNED against outputs of the reference's own
functions (tests/golden/next_rows.npz, produced by tests/golden/make_golden_next.py, which imports
models_utils.py and the trainer class verbatim).  init_level_features restates a loop that sits inline in the
trainer's run(); make_golden_next.py cuts that loop out of the reference module's source at generation time and
executes it on stand-in locals (`n{n}_x_init_ref` in the fixture), so this function is PINNED as well (bit for bit).
subgraph (cluster mini-batches) restates torch_geometric.utils.subgraph, which cannot be run here: parity unpinned.

Reference semantics worth spelling out:
  * pooling: `sums[prot_indices] += emb` / `counts[prot_indices] += 1` are numpy fancy-index updates, which apply
    ONCE per distinct index.  A protein therefore receives each DISTINCT n-gram it contains once (not once per
    occurrence); the sum runs over n-gram ids in ascending order in fp32, and `sums /= counts[:, None]` divides a
    float32 array by an int32 one (computed in float64, rounded back to float32 == the fp32 quotient).
    Proteins without any known n-gram are absent from the result.
  * labels: the successor with the largest A_out_w weight; ties are broken with random.choice, so every
    maximal successor is admissible; nodes without successors are labelled with themselves.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


def pool_proteins(protein_sequences: Sequence[Tuple[str, str]], n_val: int, ngram_map: Dict[str, int],
                  ngram_embeddings: np.ndarray) -> Tuple[List[str], np.ndarray, np.ndarray]:
    """-> (protein ids in input order, pooled [P, F] (dtype of the embeddings), valid [P] bool)."""
    num = len(protein_sequences)
    dim = ngram_embeddings.shape[1]
    sums = np.zeros((num, dim), dtype=np.float32)
    counts = np.zeros(num, dtype=np.int32)
    present: List[set] = []
    for _, seq in protein_sequences:
        ids = set()
        if len(seq) >= n_val:
            for i in range(len(seq) - n_val + 1):
                j = ngram_map.get("".join(seq[i:i + n_val]))
                if j is not None:
                    ids.add(j)
        present.append(ids)
    for p, ids in enumerate(present):               # per protein: ascending n-gram id, one add each (fp32)
        for j in sorted(ids):
            sums[p] += ngram_embeddings[j].astype(np.float32)
        counts[p] = len(ids)
    valid = counts > 0
    sums[valid] /= counts[valid, np.newaxis]
    return [pid for pid, _ in protein_sequences], sums.astype(ngram_embeddings.dtype), valid


def init_level_features(node_sequences: Sequence[str], prev_node_to_idx: Dict[str, int], prev_embeddings: np.ndarray) -> np.ndarray:
    """x[idx] = mean of the level-(n-1) embeddings of the node's prefix and suffix (those that exist), zeros if neither."""
    x = np.zeros((len(node_sequences), prev_embeddings.shape[1]), dtype=np.float32)
    for idx, s in enumerate(node_sequences):
        rows = [prev_node_to_idx.get(s[:-1]), prev_node_to_idx.get(s[1:])]
        pool = [prev_embeddings[i] for i in rows if i is not None]
        if pool:
            x[idx] = np.mean(np.array(pool, dtype=np.float32), axis=0)
    return x


def next_node_label_sets(indices: np.ndarray, values: np.ndarray, num_nodes: int) -> List[np.ndarray]:
    """Admissible next_node labels per node from the coalesced A_out_w (indices [2, E], values [E])."""
    out: List[np.ndarray] = []
    src, dst = np.asarray(indices[0]), np.asarray(indices[1])
    values = np.asarray(values)
    order = np.argsort(src, kind="stable")
    src, dst, values = src[order], dst[order], values[order]
    bounds = np.searchsorted(src, np.arange(num_nodes + 1))
    for i in range(num_nodes):
        lo, hi = bounds[i], bounds[i + 1]
        if lo == hi:
            out.append(np.array([i], dtype=np.int64))
        else:
            w = values[lo:hi]
            out.append(np.sort(dst[lo:hi][w == w.max()]).astype(np.int64))
    return out


def subgraph(subset, edge_index, edge_attr, num_nodes):
    """torch_geometric.utils.subgraph(subset, edge_index, edge_attr, relabel_nodes=True, num_nodes=N) restated from its
    published semantics (PyG is not installed here; call site protgram_directgcn_trainer.py:183-186): keep the edges
    whose two end points are in `subset`, in their original order, relabelled by node_idx[subset] = arange(len(subset)).
    PARITY PARTLY PINNED: torch_geometric (unpinned dependency of the reference, absent here) cannot be run to generate fixtures; the
    restatement is held against the known answer of PyG's own docstring example (tests/test_next_rows_oracle.py), nothing more."""
    import numpy as np
    subset = np.asarray(subset, dtype=np.int64)
    edge_index = np.asarray(edge_index)
    node_mask = np.zeros(num_nodes, dtype=bool)
    node_mask[subset] = True
    edge_mask = node_mask[edge_index[0]] & node_mask[edge_index[1]]
    node_idx = np.zeros(num_nodes, dtype=np.int64)
    node_idx[subset] = np.arange(subset.size)
    return node_idx[edge_index[:, edge_mask]], np.asarray(edge_attr)[edge_mask]
