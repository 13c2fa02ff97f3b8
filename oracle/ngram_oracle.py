"""CPU oracle for hot path A (n-gram transition-graph build) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package; the product path never does.

Fresh restatement of the reference algorithm (citations into /root/reference):

  parse_fasta            <- src/utils/data_utils.py:182-213   (DataLoader.parse_sequences)
  pad_sequences          <- src/pipeline/data_builder.py:29-35 (+ first-sequence flag :97-102)
  level_nodes            <- data_builder.py:38-42 (windows), :151-158 (distinct),
                            :164,172-173 (sorted -> id = rank)
  level_transitions      <- data_builder.py:45-54 (window i -> window i+1 id pairs),
                            :267-273 (groupby(source,target).size())
  build_level            <- the two above combined; returns what
                            aggregated_edges_n{n}.parquet + the id map hold (:281-286,:259)

Parity status: PINNED.  tests/test_oracle_golden.py checks every function here against
tests/golden/build_*.npz, which were produced by importing the reference's own modules
(tests/golden/make_golden.py).

Two implementations are kept on purpose:
  * the *_py functions walk every residue in Python exactly as the reference does
    (they are also the timed "port" CPU baseline, since that per-residue Python work
    is where the reference spends its time);
  * build_level_np is a vectorised numpy version for mid-size parity cases; it is
    itself checked against the pure-Python walk.
"""
from __future__ import annotations

import collections
from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np


def parse_fasta(path: str) -> List[Tuple[str, str]]:
    """(id, SEQUENCE) records; follows data_utils.py:182-213.

    strip every line, skip blank lines, '>' starts a record whose id is field 1 of a
    '|'-split header when present and non-empty, else the first whitespace token;
    sequence lines are upper-cased and concatenated; records with no sequence lines
    are dropped; sequence lines before any header are ignored.
    """
    out: List[Tuple[str, str]] = []
    pid = None
    parts: List[str] = []
    with open(path, "r", encoding="utf-8", errors="ignore") as fh:
        for raw in fh:
            line = raw.strip()
            if not line:
                continue
            if line[0] == ">":
                if pid and parts:
                    out.append((pid, "".join(parts)))
                hdr = line[1:]
                bar = hdr.split("|")
                pid = bar[1] if len(bar) > 1 and bar[1] else hdr.split()[0]
                parts = []
            elif pid is not None:
                parts.append(line.upper())
    if pid and parts:
        out.append((pid, "".join(parts)))
    return out


def pad_sequences(seqs: Sequence[str]) -> List[str]:
    """data_builder.py:29-35,97-102: trailing ' ' on all, leading ' ' on global sequence #0 only."""
    return [(" " if i == 0 else "") + s + " " for i, s in enumerate(seqs)]


def level_nodes_py(padded: Iterable[str], n: int) -> List[str]:
    """Sorted distinct n-grams; list position is the node id (data_builder.py:38-42,164,172-173)."""
    seen = set()
    for p in padded:
        if len(p) >= n:
            for i in range(len(p) - n + 1):
                seen.add(p[i:i + n])
    return sorted(seen)


def level_transitions_py(padded: Iterable[str], n: int, node_id: Dict[str, int]) -> Dict[Tuple[int, int], int]:
    """Multiset of (id(window i), id(window i+1)) -> count (data_builder.py:45-54,267-273)."""
    cnt: Dict[Tuple[int, int], int] = collections.Counter()
    for p in padded:
        if len(p) >= n + 1:
            for i in range(len(p) - n):
                s = node_id.get(p[i:i + n])
                t = node_id.get(p[i + 1:i + 1 + n])
                if s is not None and t is not None:
                    cnt[(s, t)] += 1
    return cnt


def build_level_py(padded: Sequence[str], n: int):
    """-> (nodes: list[str], src int64[E], dst int64[E], count int64[E]) sorted by (src, dst)."""
    nodes = level_nodes_py(padded, n)
    node_id = {g: i for i, g in enumerate(nodes)}
    cnt = level_transitions_py(padded, n, node_id)
    keys = sorted(cnt)
    src = np.fromiter((k[0] for k in keys), dtype=np.int64, count=len(keys))
    dst = np.fromiter((k[1] for k in keys), dtype=np.int64, count=len(keys))
    w = np.fromiter((cnt[k] for k in keys), dtype=np.int64, count=len(keys))
    return nodes, src, dst, w


def _windows_codes(padded: Sequence[str], m: int) -> np.ndarray:
    """All length-m windows of every padded sequence as base-256 integer codes (uint64, m<=8)."""
    chunks = []
    for p in padded:
        if len(p) >= m:
            b = np.frombuffer(p.encode("latin-1"), dtype=np.uint8).astype(np.uint64)
            L = len(b) - m + 1
            code = np.zeros(L, dtype=np.uint64)
            for k in range(m):
                code = code * np.uint64(256) + b[k:k + L]
            chunks.append(code)
    return np.concatenate(chunks) if chunks else np.zeros(0, dtype=np.uint64)


def _decode(code: int, m: int) -> str:
    return bytes((code >> (8 * (m - 1 - k))) & 0xFF for k in range(m)).decode("latin-1")


def build_level_np(padded: Sequence[str], n: int):
    """Vectorised build_level (same return contract).  Requires latin-1 encodable text, n <= 7."""
    assert 1 <= n <= 7
    ncodes = np.unique(_windows_codes(padded, n))  # base-256 code order == byte-lexicographic order
    nodes = [_decode(int(c), n) for c in ncodes]
    ecodes, w = np.unique(_windows_codes(padded, n + 1), return_counts=True)
    src_code = ecodes >> np.uint64(8)
    dst_code = ecodes & np.uint64((1 << (8 * n)) - 1)
    src = np.searchsorted(ncodes, src_code).astype(np.int64)
    dst = np.searchsorted(ncodes, dst_code).astype(np.int64)
    order = np.lexsort((dst, src))
    return nodes, src[order], dst[order], w.astype(np.int64)[order]


def build_all_levels(seqs: Sequence[str], n_max: int, impl: str = "py"):
    """GraphBuilder.run() phases 1+2 without the file hand-offs: {n: (nodes, src, dst, count)}."""
    padded = pad_sequences(seqs)
    fn = build_level_py if impl == "py" else build_level_np
    return {n: fn(padded, n) for n in range(1, n_max + 1)}


# ---------------------------------------------------------------------------------------------
# synthetic corpus shared by tests and bench: deterministic, shard-independent.
# The product generates the same corpus on the device (csrc/synth.cu); this is the CPU twin.
# ---------------------------------------------------------------------------------------------
AA = "ACDEFGHIKLMNPQRSTVWY"
# cumulative UniProt-like background frequencies scaled to 2^16 (same table as csrc/synth.cu)
AA_CUM16 = np.array([5408, 6308, 9885, 14308, 16838, 21473, 22961, 26847, 30662, 37140, 38723,
                     41383, 44486, 47063, 50689, 54995, 58502, 63002, 63720, 65536], dtype=np.uint32)


def _mix32(x: np.ndarray) -> np.ndarray:
    """lowbias32-style integer hash on uint32 arrays (wraps mod 2^32)."""
    x = x.astype(np.uint64)
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7FEB352D)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x846CA68B)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    return x.astype(np.uint32)


def synth_residues(first_seq: int, nseq: int, seq_len: int, seed: int = 42) -> np.ndarray:
    """uint8 [nseq, seq_len] residues; residue (s, j) depends only on (seed, s, j)."""
    s = (np.arange(first_seq, first_seq + nseq, dtype=np.uint64)[:, None] * np.uint64(seq_len)
         + np.arange(seq_len, dtype=np.uint64)[None, :])
    lo = (s & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    hi = (s >> np.uint64(32)).astype(np.uint32)
    h = _mix32(lo ^ _mix32(hi ^ np.uint32(seed)) ^ np.uint32(0x9E3779B9))
    u = (h >> np.uint32(16)).astype(np.uint32)  # 16 uniform bits
    idx = np.searchsorted(AA_CUM16, u, side="right")
    return np.frombuffer(AA.encode(), dtype=np.uint8)[idx]


def synth_sequences(first_seq: int, nseq: int, seq_len: int, seed: int = 42) -> List[str]:
    r = synth_residues(first_seq, nseq, seq_len, seed)
    return [row.tobytes().decode("ascii") for row in r]
