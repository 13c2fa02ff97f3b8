"""EmbeddingProcessor -- the methods of the reference's src/utils/models_utils.py that sit on or next to
hot path B: l2_normalize_torch (:139-147, last op of ProtGramDirectGCN.forward),
extract_gcn_node_embeddings (:265-273, eval-mode forward -> numpy) and the protein-level pooling
pool_ngram_embeddings_for_protein_fast (:210-262)."""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from .. import _native as nat


def encode_ngrams(ngrams: Sequence[str], n: int, alphabet: str = None) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """n-gram strings -> packed base-sigma codes over the byte-sorted alphabet of the strings (most significant
    digit = first character, so code order == string order).  -> (codes int64, symbols uint8[sigma], rank uint8[256],
    255 = byte outside the alphabet).  Non-ASCII n-grams raise ValueError (the byte kernels cannot window them)."""
    try:
        raw = np.frombuffer("".join(ngrams).encode("ascii"), dtype=np.uint8)
    except UnicodeEncodeError as exc:
        raise ValueError("non-ASCII n-grams are unsupported on the CUDA path") from exc
    if any(len(s) != n for s in ngrams):
        raise ValueError(f"every n-gram must have exactly {n} characters")
    chars = raw.reshape(len(ngrams), n) if len(ngrams) else np.zeros((0, n), dtype=np.uint8)
    symbols = np.unique(np.frombuffer(alphabet.encode("ascii"), dtype=np.uint8)) if alphabet is not None else np.unique(chars)
    rank = np.full(256, 255, dtype=np.uint8)
    rank[symbols] = np.arange(symbols.size, dtype=np.uint8)
    codes = np.zeros(len(ngrams), dtype=np.int64)
    for k in range(n):
        codes = codes * int(symbols.size) + rank[chars[:, k]].astype(np.int64)
    return codes, symbols, rank


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
    def pool_ngram_embeddings_for_protein_fast(protein_sequences: List[Tuple[str, str]], n_val: int, ngram_map: Dict[str, int],
                                               ngram_embeddings: np.ndarray) -> Dict[str, np.ndarray]:
        """Reference :210-262, same signature and result: {protein id: mean embedding of the DISTINCT known n-grams of
        the protein}; proteins without a known n-gram are absent.  One CTA per protein on the GPU instead of a Python
        loop over every residue (csrc/next.cu).  Bit-identical to the reference when the ids in `ngram_map` are the
        ranks of the sorted n-grams (what GraphBuilder produces); for any other id assignment the fp32 summation
        order differs (ascending code instead of ascending id)."""
        if not protein_sequences:
            return {}
        nat.require_cuda()
        dev = nat.current_device()
        names = list(ngram_map.keys())
        ids = np.fromiter((ngram_map[k] for k in names), dtype=np.int64, count=len(names))
        codes, symbols, rank = encode_ngrams(names, n_val)
        sigma = int(symbols.size)
        table = np.full(max(sigma, 1) ** n_val, -1, dtype=np.int32)
        table[codes] = ids.astype(np.int32)
        try:
            blob = "".join(seq if isinstance(seq, str) else "".join(seq) for _, seq in protein_sequences).encode("ascii")
        except UnicodeEncodeError as exc:
            raise ValueError("non-ASCII protein sequences are unsupported on the CUDA path") from exc
        offsets = np.zeros(len(protein_sequences) + 1, dtype=np.int64)
        np.cumsum([len(seq) for _, seq in protein_sequences], out=offsets[1:])
        emb = torch.from_numpy(np.ascontiguousarray(ngram_embeddings, dtype=np.float32)).to(dev)
        d_seq = torch.from_numpy(np.frombuffer(blob, dtype=np.uint8).copy() if blob else np.zeros(1, np.uint8)).to(dev)
        d_off, d_rank, d_tab = (torch.from_numpy(a).to(dev) for a in (offsets, rank, table))
        num, dim = len(protein_sequences), int(emb.shape[1])
        out = torch.empty((num, dim), dtype=torch.float32, device=dev)
        valid = torch.empty(num, dtype=torch.uint8, device=dev)
        nat.call("pg_pool_proteins", nat.ptr(d_seq), nat.ptr(d_off), num, int(n_val), nat.ptr(d_rank), sigma, nat.ptr(d_tab),
                 nat.ptr(emb), emb.stride(0), dim, nat.ptr(out), out.stride(0), nat.ptr(valid), nat.stream_ptr())
        pooled = out.cpu().numpy().astype(ngram_embeddings.dtype, copy=False)
        keep = valid.cpu().numpy().astype(bool)
        return {protein_sequences[i][0]: pooled[i] for i in np.nonzero(keep)[0]}

    @staticmethod
    def extract_gcn_node_embeddings(model, data, device: torch.device) -> np.ndarray:
        model.eval()
        model.to(device)
        data = data.to(device)
        with torch.no_grad():
            _, embeddings = model(data=data)
        return embeddings.cpu().numpy()
