"""The per-level preparation steps of the reference trainer that loop over every node / every cluster in Python
(src/pipeline/protgram_directgcn_trainer.py), as kernels: next-node labels, the feature
hand-off between levels, and the cluster mini-batch extraction.  Same results, same argument meaning; the training
loop itself (and the METIS / Louvain partitioning that produces the clusters) stays the reference's."""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from .. import _native as nat
from .models_utils import encode_ngrams


def generate_next_node_labels(graph) -> Tuple[torch.Tensor, int]:
    """Reference `_generate_next_node_labels` (:222-237): label of a node = its successor with the largest A_out_w
    weight (the reference breaks ties with random.choice; here: the first maximal successor), a node without
    successors is labelled with itself.  -> (labels int64 [N] on the CPU like the reference, num_classes = N)."""
    n = graph.number_of_nodes
    if n == 0:
        return torch.empty(0, dtype=torch.long), 1
    nat.require_cuda()
    dev = nat.current_device()
    a = graph.A_out_w.coalesce()
    idx, val = a.indices().to(dev), a.values().to(dev, dtype=torch.float32).contiguous()
    src, dst = idx[0].contiguous(), idx[1].contiguous()
    rowptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    st = nat.stream_ptr()
    nat.call("pg_rowptr_from_sorted", nat.ptr(src), src.numel(), n, nat.ptr(rowptr), st)
    labels = torch.empty(n, dtype=torch.int64, device=dev)
    nat.call("pg_next_node_labels", nat.ptr(rowptr), nat.ptr(dst), nat.ptr(val), n, nat.ptr(labels), st)
    return labels.cpu(), n


def init_level_features(graph, prev_node_to_idx: Dict[str, int], prev_embeddings: np.ndarray, device=None) -> torch.Tensor:
    """Reference :312-330: x[idx] = mean of the level-(n-1) embeddings of the node's prefix and suffix (those present
    in `prev_node_to_idx`), zeros if neither.  -> float32 [N, F] on `device` (default: the current CUDA device)."""
    nat.require_cuda()
    dev = torch.device(device) if device is not None else nat.current_device()
    nodes = list(graph.node_sequences)
    n = len(nodes[0]) if nodes else 0
    f = int(prev_embeddings.shape[1])
    x = torch.zeros((len(nodes), f), dtype=torch.float32, device=dev)
    if not nodes or n < 2 or not prev_node_to_idx:
        return x
    prev_names = list(prev_node_to_idx.keys())
    alphabet = "".join(sorted(set("".join(nodes)) | set("".join(prev_names))))
    codes, symbols, _ = encode_ngrams(nodes, n, alphabet)
    prev_codes, _, _ = encode_ngrams(prev_names, n - 1, alphabet)
    order = np.argsort(prev_codes, kind="stable")
    prev_rows = np.fromiter((prev_node_to_idx[k] for k in prev_names), dtype=np.int64, count=len(prev_names))[order]
    emb_sorted = np.ascontiguousarray(np.asarray(prev_embeddings, dtype=np.float32)[prev_rows])   # row k <-> k-th smallest code
    d_code, d_prev = torch.from_numpy(codes).to(dev), torch.from_numpy(prev_codes[order]).to(dev)
    d_emb = torch.from_numpy(emb_sorted).to(dev)
    nat.call("pg_ngram_feature_init", nat.ptr(d_code), len(nodes), nat.ptr(d_prev), len(prev_names), int(symbols.size), n,
             nat.ptr(d_emb), d_emb.stride(0), f, nat.ptr(x), x.stride(0), nat.stream_ptr())
    return x


def _device_csr(graph, dev):
    """(rowptr int64[N+1], col int32[P], (val_in, val_out, val_und)) of the shared pattern on `dev`."""
    side = getattr(graph, "_pg_device", None)
    if side is not None and side["col"].device == dev:
        return side["rowptr"], side["col"], (side["val_in"], side["val_out"], side["val_und"])
    m_in, m_out, m_un = (t.coalesce() for t in (graph.mathcal_A_in, graph.mathcal_A_out, graph.A_undirected_norm_sparse))
    idx = m_in.indices().to(dev)
    if not (torch.equal(m_out.indices().to(dev), idx) and torch.equal(m_un.indices().to(dev), idx)):
        raise ValueError("cluster extraction needs the three propagation matrices on one shared pattern (reference-built graphs have it)")
    n = graph.number_of_nodes
    rows = idx[0].contiguous()
    rowptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    nat.call("pg_rowptr_from_sorted", nat.ptr(rows), rows.numel(), n, nat.ptr(rowptr), nat.stream_ptr())
    vals = tuple(t.values().to(dev, dtype=torch.float32).contiguous() for t in (m_in, m_out, m_un))
    return rowptr, idx[1].to(torch.int32).contiguous(), vals


def create_clustered_subgraphs(graph, cluster_list: Sequence[Sequence[int]], full_data, device=None) -> List:
    """Reference `_create_clustered_subgraphs` from :178 on (the loop over clusters; the partitioning above it is not
    on the hot path): per cluster the induced subgraph of all three propagation matrices, relabelled by position in
    the cluster, `x` / `y` rows of the cluster, `original_indices`.  One pass over the shared-pattern CSR per cluster
    instead of three `torch_geometric.utils.subgraph` calls over the full edge lists; for ascending clusters (what
    the reference's partition dict yields) the edge order is the reference's and the layer gets the sub-CSR as is."""
    from .protgram_directgcn import Data, register_symmetric_structure
    nat.require_cuda()
    dev = torch.device(device) if device is not None else nat.current_device()
    n = graph.number_of_nodes
    rowptr, col, vals = _device_csr(graph, dev)
    new_id = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    x_full = full_data.x.to(dev)
    y_full = getattr(full_data, "y", None)
    st = nat.stream_ptr()
    out = []
    for cluster_nodes in cluster_list:
        sub = torch.as_tensor(cluster_nodes, dtype=torch.long).to(dev)
        k = int(sub.numel())
        sub_rowptr = torch.empty(k + 1, dtype=torch.int64, device=dev)
        ws = nat.workspace(nat.query("pg_subgraph_ws_bytes", k), dev)
        nat.call("pg_subgraph_sizes", nat.ptr(rowptr), nat.ptr(col), n, nat.ptr(sub), k, nat.ptr(new_id), nat.ptr(sub_rowptr),
                 nat.ptr(ws), ws.numel(), st)
        p = int(sub_rowptr[k].item())
        sub_col = torch.empty(p, dtype=torch.int32, device=dev)
        w = [torch.empty(p, dtype=torch.float32, device=dev) for _ in range(3)]
        ei = torch.empty((2, p), dtype=torch.int64, device=dev)
        nat.call("pg_subgraph_fill", nat.ptr(rowptr), nat.ptr(col), nat.ptr(vals[0]), nat.ptr(vals[1]), nat.ptr(vals[2]), n, nat.ptr(sub), k,
                 nat.ptr(new_id), nat.ptr(sub_rowptr), nat.ptr(sub_col), nat.ptr(w[0]), nat.ptr(w[1]), nat.ptr(w[2]), nat.ptr(ei[0]),
                 nat.ptr(ei[1]), st)
        if k > 1 and bool((sub[1:] > sub[:-1]).all()):      # ascending cluster: sorted symmetric sub-CSR, usable as is
            register_symmetric_structure(ei, tuple(w), k, sub_rowptr, sub_col, static=False)
        y = y_full[sub.to(y_full.device)] if (y_full is not None and y_full.numel() > 0) else torch.empty(0)
        out.append(Data(x=x_full[sub], y=y, edge_index_in=ei, edge_weight_in=w[0], edge_index_out=ei, edge_weight_out=w[1],
                        edge_index_undirected_norm=ei, edge_weight_undirected_norm=w[2], original_indices=sub))
    return out
