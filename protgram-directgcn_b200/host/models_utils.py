"""EmbeddingProcessor -- the two methods of the reference's src/utils/models_utils.py that sit on
hot path B: l2_normalize_torch (:139-147, last op of ProtGramDirectGCN.forward) and
extract_gcn_node_embeddings (:265-273, eval-mode forward -> numpy)."""
from __future__ import annotations

import numpy as np
import torch

from .. import _native as nat


class EmbeddingProcessor:
    @staticmethod
    def l2_normalize_torch(embeddings: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
        if embeddings.ndim not in (1, 2):
            raise ValueError(f"Unsupported tensor ndim for L2 normalization: {embeddings.ndim}. Expected 1 or 2.")
        needs_graph = torch.is_grad_enabled() and embeddings.requires_grad
        if embeddings.is_cuda and embeddings.ndim == 2 and embeddings.dtype == torch.float32 and not needs_graph:
            h = embeddings.contiguous()
            out = torch.empty_like(h)
            nat.call("pg_l2_normalize_rows", nat.ptr(h), h.stride(0), h.shape[0], h.shape[1], float(eps), nat.ptr(out),
                     out.stride(0), nat.stream_ptr())
            return out
        # differentiable / 1-D form (plain torch ops, autograd-visible)
        norm = torch.norm(embeddings, p=2, dim=embeddings.ndim - 1, keepdim=True)
        return embeddings / (norm + eps)

    @staticmethod
    def extract_gcn_node_embeddings(model, data, device: torch.device) -> np.ndarray:
        model.eval()
        model.to(device)
        data = data.to(device)
        with torch.no_grad():
            _, embeddings = model(data=data)
        return embeddings.cpu().numpy()
