"""CPU oracle for the graph normalisation half of hot path A -- TEST INFRASTRUCTURE ONLY.

Restates, in numpy float32, what DirectedNgramGraph.__init__ computes from the aggregated
edge table (citations into /root/reference/src/utils/graph_utils.py):

  raw_adjacency            <- :109-112 (int64 ids, float32(count)), :140-158 (coalesced A_out_w,
                              A_in_w = A_out_w^T coalesced)
  undirected_normalized    <- :160-196 (unique pairs -> symmetrise -> unique -> append N self
                              loops unconditionally -> deg = occurrences of each column ->
                              deg^-1/2[row] * 1 * deg^-1/2[col] -> coalesce (duplicates summed))
  propagation_matrix       <- :198-273 (D^-1 A; 0.5*(An^2 + (An^2)^T); sqrt(. + eps); + I)

All results are (row int64[P], col int64[P], val float32[P]) in row-major sorted (coalesced)
order, i.e. exactly `.indices()` / `.values()` of the reference's sparse COO tensors.

Parity status: PINNED against tests/golden/build_*.npz (reference-generated).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def _coalesce(row, col, val, n):
    """Sort row-major and sum duplicates sequentially in float32 (torch coalesce on CPU)."""
    row = np.asarray(row, dtype=np.int64)
    col = np.asarray(col, dtype=np.int64)
    val = np.asarray(val, dtype=F32)
    if row.size == 0:
        return row, col, val
    key = row * np.int64(n) + col
    order = np.argsort(key, kind="stable")
    key, val = key[order], val[order]
    first = np.ones(key.size, dtype=bool)
    first[1:] = key[1:] != key[:-1]
    seg = np.cumsum(first) - 1
    out = np.zeros(int(seg[-1]) + 1, dtype=F32)
    np.add.at(out, seg, val)  # sequential float32 accumulation in sorted order
    ukey = key[first]
    return ukey // n, ukey % n, out


def raw_adjacency(src, dst, count, n):
    """-> (A_out_w, A_in_w) each as (row, col, val)."""
    w = np.asarray(count).astype(F32)  # int64 -> float32 round-to-nearest-even (:112)
    a_out = _coalesce(src, dst, w, n)
    a_in = _coalesce(a_out[1], a_out[0], a_out[2], n)
    return a_out, a_in


def undirected_normalized(src, dst, n):
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    pairs = np.unique(np.stack([src, dst], axis=1), axis=0) if src.size else np.zeros((0, 2), np.int64)
    sym = np.concatenate([pairs, pairs[:, ::-1]], axis=0)
    sym = np.unique(sym, axis=0) if sym.size else sym
    loops = np.arange(n, dtype=np.int64)
    row = np.concatenate([sym[:, 0], loops])
    col = np.concatenate([sym[:, 1], loops])
    deg = np.bincount(col, minlength=n).astype(F32)
    with np.errstate(divide="ignore"):
        dis = (F32(1.0) / np.sqrt(deg)).astype(F32)
    dis[np.isinf(dis)] = 0
    val = (dis[row] * F32(1.0)) * dis[col]
    return _coalesce(row, col, val, n)


def propagation_matrix(a, n, eps=1e-9):
    """a = (row, col, val) of a coalesced weighted adjacency (A_out_w or A_in_w)."""
    row, col, val = a
    if n == 0 or row.size == 0:
        z = np.zeros(0, np.int64)
        return z, z.copy(), np.zeros(0, F32)
    rs = np.zeros(n, dtype=F32)
    np.add.at(rs, row, val)
    dinv = np.zeros(n, dtype=F32)
    nz = rs != 0
    dinv[nz] = F32(1.0) / rs[nz]
    an = (val * dinv[row]).astype(F32)
    sq = (an * an).astype(F32)
    r2, c2, v2 = _coalesce(np.concatenate([row, col]), np.concatenate([col, row]),
                           np.concatenate([sq, sq]), n)  # An^2 + (An^2)^T
    v2 = (v2 * F32(0.5)).astype(F32)
    base = np.sqrt((v2 + F32(eps)).astype(F32)).astype(F32)
    loops = np.arange(n, dtype=np.int64)
    return _coalesce(np.concatenate([r2, loops]), np.concatenate([c2, loops]),
                     np.concatenate([base, np.ones(n, dtype=F32)]), n)


def normalise_all(src, dst, count, n, eps=1e-9):
    """Everything DirectedNgramGraph holds, as a dict name -> (row, col, val)."""
    a_out, a_in = raw_adjacency(src, dst, count, n)
    return {
        "A_out_w": a_out,
        "A_in_w": a_in,
        "A_undirected_norm_sparse": undirected_normalized(src, dst, n),
        "mathcal_A_out": propagation_matrix(a_out, n, eps),
        "mathcal_A_in": propagation_matrix(a_in, n, eps),
    }
