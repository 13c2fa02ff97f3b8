"""The two per-level preparation steps of the reference trainer that loop over every node in Python
(src/pipeline/protgram_directgcn_trainer.py), as kernels over the packed n-gram codes.
Same results, same argument meaning; the training loop itself stays the reference's."""
from __future__ import annotations

from typing import Dict, Tuple

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
