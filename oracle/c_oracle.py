"""ctypes wrapper around oracle/ngram_count.c -- TEST INFRASTRUCTURE ONLY.

Also holds the oracle's own copy of the corpus-buffer packing (include/pgb200.h "corpus buffer")
so tests can feed the CUDA path and the C oracle the same bytes.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import List, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
SEP = 0xFF


def build() -> str:
    so = os.path.join(_HERE, "_build", "liboracle.so")
    src = os.path.join(_HERE, "ngram_count.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.oracle_ngram_count.restype = ctypes.c_int
        _LIB.oracle_byte_presence.restype = ctypes.c_int
    return _LIB


def pack_corpus(seqs: Sequence[str], first_is_global_first: bool = True) -> np.ndarray:
    """padded sequences (data_builder.py:29-35) each followed by the 0xFF separator."""
    parts: List[bytes] = []
    for i, s in enumerate(seqs):
        b = s.encode("ascii")
        parts.append((b" " if (i == 0 and first_is_global_first) else b"") + b + b" \xff")
    return np.frombuffer(b"".join(parts), dtype=np.uint8).copy()


def alphabet(buf: np.ndarray):
    pres = np.zeros(256, dtype=np.uint8)
    lib().oracle_byte_presence(buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(buf.size),
                               pres.ctypes.data_as(ctypes.c_void_p))
    symbols = np.nonzero(pres)[0].astype(np.uint8)
    rank = np.zeros(256, dtype=np.uint8)
    rank[symbols] = np.arange(symbols.size, dtype=np.uint8)
    return symbols, rank


def count_level(buf: np.ndarray, n: int, rank: np.ndarray, sigma: int):
    bins = np.zeros(sigma ** (n + 1), dtype=np.uint64)
    present = np.zeros(sigma ** n, dtype=np.uint8)
    lib().oracle_ngram_count(buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(buf.size),
                             ctypes.c_int(n), rank.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(sigma),
                             bins.ctypes.data_as(ctypes.c_void_p), present.ctypes.data_as(ctypes.c_void_p))
    return bins, present


def bins_to_graph(bins: np.ndarray, present: np.ndarray, symbols: np.ndarray, n: int):
    """dense tables -> (nodes, src, dst, count) in the ngram_oracle.build_level contract."""
    sigma = symbols.size
    node_codes = np.nonzero(present)[0]
    ids = np.cumsum(present.astype(np.int64)) - 1
    nodes = []
    for c in node_codes:
        c = int(c)
        chars = []
        for _ in range(n):
            chars.append(int(symbols[c % sigma]))
            c //= sigma
        nodes.append(bytes(reversed(chars)).decode("ascii"))
    ecodes = np.nonzero(bins)[0]
    src = ids[ecodes // sigma]
    dst = ids[ecodes % (sigma ** n)]
    return nodes, src.astype(np.int64), dst.astype(np.int64), bins[ecodes].astype(np.int64)
