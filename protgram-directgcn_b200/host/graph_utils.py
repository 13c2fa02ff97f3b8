"""Graph / DirectedNgramGraph -- same public surface as the reference's src/utils/graph_utils.py
(:19-125): constructor signature, attribute names, coalesced fp32 sparse-COO matrices with int64
indices.  The matrix construction (reference :140-287) runs on the GPU through libpgb200.so:
one radix sort of 2E+N tagged keys + two per-item passes (csrc/graph.cu) instead of ~14
sparse coalesces.  There is no CPU fallback: building matrices without a CUDA device raises.
"""
from __future__ import annotations

import os
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

from .. import _native as nat


MATRIX_ATTRS = ("A_out_w", "A_in_w", "A_undirected_norm_sparse", "mathcal_A_out", "mathcal_A_in")


def _matrix_property(name):
    key = "_m_" + name

    def get(self):
        self.wait_ready()
        try:
            return self.__dict__[key]
        except KeyError:
            raise AttributeError(name) from None

    def set_(self, value):
        self.__dict__[key] = value

    return property(get, set_)


class _LazyMaps:
    """idx_to_node / node_to_idx are views derived from node_sequences; they are built on first
    access (an 8k-entry dict costs more host time than the whole GPU graph build) and always
    written into pickles, so a pickled graph has the reference's attribute layout."""

    @property
    def node_sequences(self):
        """id -> n-gram list.  GraphBuilder hands over the packed node codes still on their way from
        the device (corpus.LazyNodeNames); they are decoded on first access, i.e. under the GPU work
        that was enqueued in the meantime."""
        d = self.__dict__
        lazy = d.get("_lazy_names")
        if lazy is not None:
            d["_node_sequences"] = lazy.resolve()
            d["_lazy_names"] = None
        return d.setdefault("_node_sequences", [])

    @node_sequences.setter
    def node_sequences(self, value):
        self.__dict__["_node_sequences"] = value
        self.__dict__["_lazy_names"] = None

    def wait_ready(self):
        """Block until the host copies of the matrices (started asynchronously by the GPU build on a side
        stream) have landed.  Every access to a matrix attribute goes through here."""
        ev = self.__dict__.get("_ready_event")
        if ev is not None:
            ev.synchronize()
            self.__dict__["_ready_event"] = None

    def _maps(self):
        d = self.__dict__
        if d.get("_idx_to_node") is None:
            names = self.node_sequences
            d["_idx_to_node"] = dict(zip(range(len(names)), names))
            d["_node_to_idx"] = dict(zip(names, range(len(names))))
        return d["_idx_to_node"], d["_node_to_idx"]

    @property
    def idx_to_node(self):
        return self._maps()[0]

    @idx_to_node.setter
    def idx_to_node(self, value):
        self._maps()
        self.__dict__["_idx_to_node"] = value

    @property
    def node_to_idx(self):
        return self._maps()[1]

    @node_to_idx.setter
    def node_to_idx(self, value):
        self._maps()
        self.__dict__["_node_to_idx"] = value

    def __getstate__(self):
        state = dict(self.__dict__)
        i2n, n2i = self._maps()
        state.pop("_idx_to_node", None)
        state.pop("_node_to_idx", None)
        self.wait_ready()
        state.pop("_ready_event", None)
        for m in MATRIX_ATTRS:
            if "_m_" + m in state:
                state[m] = state.pop("_m_" + m)
        state.pop("_lazy_names", None)
        state.pop("_node_sequences", None)
        state["node_sequences"] = self.node_sequences
        state["idx_to_node"], state["node_to_idx"] = i2n, n2i
        state.pop("_pg_device", None)  # device-side CSR sidecar never enters the pickle
        state.pop("_pg_sidecar", None)  # nor the host-side one loaded from <path>.csr.npz
        return state

    def __setstate__(self, state):
        state = dict(state)
        self.__dict__["_idx_to_node"] = state.pop("idx_to_node", None)
        self.__dict__["_node_to_idx"] = state.pop("node_to_idx", None)
        self.__dict__["_node_sequences"] = state.pop("node_sequences", [])
        self.__dict__["_lazy_names"] = None
        for m in MATRIX_ATTRS:
            if m in state:
                self.__dict__["_m_" + m] = state.pop(m)
        self.__dict__.update(state)


for _m in MATRIX_ATTRS:   # plain attributes in the reference; here they first wait for the async host copy
    setattr(_LazyMaps, _m, _matrix_property(_m))


class Graph(_LazyMaps):
    """Base container: id <-> n-gram maps (reference graph_utils.py:19-87)."""

    def __init__(self, nodes: Dict[int, Any], edges: List[Tuple]):
        self.idx_to_node_map_from_constructor = nodes if nodes is not None else {}
        self.original_edges = edges if edges is not None else []
        self.number_of_nodes: int = 0
        self.node_sequences: List[Any] = []
        self.edges: List[Tuple] = []
        self.number_of_edges: int = 0
        self._process_constructor_inputs()

    def _process_constructor_inputs(self):
        node_map = self.idx_to_node_map_from_constructor
        if isinstance(node_map, list):
            # internal fast path: GraphBuilder hands over the decoded names in id order
            self.idx_to_node_map_from_constructor = {}
            self.number_of_nodes = len(node_map)
            self.node_sequences = node_map
            return
        if hasattr(node_map, "resolve"):
            # ... or the packed codes still in flight from the device (decoded on first access)
            self.idx_to_node_map_from_constructor = {}
            self.number_of_nodes = len(node_map)
            self.__dict__["_lazy_names"] = node_map
            return
        if not node_map and not self.original_edges:
            return
        top = -1
        for key in node_map.keys():
            if isinstance(key, (int, np.integer)) and key >= 0:
                top = max(top, int(key))
        for e in self.original_edges:
            if len(e) >= 2 and isinstance(e[0], (int, np.integer)) and isinstance(e[1], (int, np.integer)):
                top = max(top, int(e[0]), int(e[1]))
        self.number_of_nodes = top + 1
        names = [node_map.get(i) for i in range(self.number_of_nodes)]
        self.node_sequences = [str(nm) if nm is not None else f"__NODE_{i}__" for i, nm in enumerate(names)]
        self.edges = self.original_edges
        self.number_of_edges = len(self.edges)


def _coo(indices: torch.Tensor, values: torch.Tensor, n: int) -> torch.Tensor:
    return torch.sparse_coo_tensor(indices, values, (n, n), is_coalesced=True)


def _empty_coo(n: int, device="cpu") -> torch.Tensor:
    return _coo(torch.empty((2, 0), dtype=torch.long, device=device),
                torch.empty(0, dtype=torch.float32, device=device), n)


_D2H_STREAMS: Dict[int, "torch.cuda.Stream"] = {}


def _async_to_host(tensors):
    """Device tensors -> pinned host tensors (torch's caching host allocator recycles the blocks), copied on
    a side stream so the caller's stream is free to run the next kernels.  -> (host tensors, done event)."""
    dev = tensors[0].device
    side = _D2H_STREAMS.get(dev.index)
    if side is None:
        side = _D2H_STREAMS[dev.index] = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    out = []
    with torch.cuda.stream(side):
        for t in tensors:
            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            h.copy_(t, non_blocking=True)
            t.record_stream(side)
            out.append(h)
        ev = torch.cuda.Event()
        ev.record(side)
    return out, ev


def device_coalesce(src: torch.Tensor, dst: torch.Tensor, w: torch.Tensor, n: int):
    """Arbitrary edge table -> row-major sorted, duplicates summed (reference :154 `.coalesce()`)."""
    nnz = src.numel()
    dev = src.device
    so, do, wo = torch.empty_like(src), torch.empty_like(dst), torch.empty_like(w)
    sizes = torch.zeros(1, dtype=torch.int64, device=dev)
    ws = nat.workspace(nat.query("pg_coo_coalesce_ws_bytes", nnz), dev)
    nat.call("pg_coo_coalesce", nat.ptr(src), nat.ptr(dst), nat.ptr(w), nnz, n, nat.ptr(so), nat.ptr(do), nat.ptr(wo),
             nat.ptr(sizes), nat.ptr(ws), ws.numel(), nat.stream_ptr())
    e = int(sizes.item())
    return so[:e], do[:e], wo[:e]


def device_normalize(src: torch.Tensor, dst: torch.Tensor, w: torch.Tensor, n: int, eps: float):
    """Coalesced A_out_w (device, sorted unique) -> dict with A_in_w COO, shared CSR pattern and
    the three value arrays.  Replaces reference :158-287."""
    nnz = src.numel()
    dev = src.device
    ws = nat.workspace(nat.query("pg_normalize_ws_bytes", nnz, n), dev)
    sizes = torch.zeros(1, dtype=torch.int64, device=dev)
    st = nat.stream_ptr()
    nat.call("pg_normalize_sizes", nat.ptr(src), nat.ptr(dst), nat.ptr(w), nnz, n, nat.ptr(sizes), nat.ptr(ws), ws.numel(), st)
    p = int(sizes.item())
    in_src = torch.empty(nnz, dtype=torch.int64, device=dev)
    in_dst = torch.empty(nnz, dtype=torch.int64, device=dev)
    in_w = torch.empty(nnz, dtype=torch.float32, device=dev)
    rowptr = torch.empty(n + 1, dtype=torch.int64, device=dev)
    col = torch.empty(p, dtype=torch.int32, device=dev)
    v_out = torch.empty(p, dtype=torch.float32, device=dev)
    v_in = torch.empty(p, dtype=torch.float32, device=dev)
    v_und = torch.empty(p, dtype=torch.float32, device=dev)
    nat.call("pg_normalize_fill", nat.ptr(src), nat.ptr(dst), nat.ptr(w), nnz, n, float(eps), p, nat.ptr(in_src),
             nat.ptr(in_dst), nat.ptr(in_w), nat.ptr(rowptr), nat.ptr(col), nat.ptr(v_out), nat.ptr(v_in), nat.ptr(v_und),
             nat.ptr(ws), ws.numel(), st)
    return {"in_src": in_src, "in_dst": in_dst, "in_w": in_w, "rowptr": rowptr, "col": col,
            "val_out": v_out, "val_in": v_in, "val_und": v_und, "pattern_nnz": p}


def csr_to_coo_indices(rowptr: torch.Tensor, col: torch.Tensor, n: int) -> torch.Tensor:
    """int32 CSR -> the reference's int64 [2, P] COO index layout."""
    p = col.numel()
    idx = torch.empty((2, p), dtype=torch.int64, device=col.device)
    nat.call("pg_coo_from_csr", nat.ptr(rowptr), nat.ptr(col), n, p, nat.ptr(idx[0]), nat.ptr(idx[1]), nat.stream_ptr())
    return idx


class DirectedNgramGraph(Graph):
    """Drop-in for reference graph_utils.py:90-287.  Attributes: A_out_w, A_in_w,
    A_undirected_norm_sparse, mathcal_A_out, mathcal_A_in (coalesced sparse COO, fp32, int64
    indices, on the CPU after construction exactly like the reference), node_to_idx, idx_to_node,
    node_sequences, number_of_nodes, number_of_edges, n_value, epsilon_propagation."""

    def __init__(self, nodes: Dict[int, Any], edge_file_path: Optional[str] = None,
                 epsilon_propagation: float = 1e-9, n_value: Optional[int] = None):
        super().__init__(nodes=nodes, edges=[])
        self.epsilon_propagation = epsilon_propagation
        self.n_value: Optional[int] = n_value
        if self.number_of_nodes > 0 and edge_file_path and os.path.exists(edge_file_path):
            print(f"    Loading edges from {os.path.basename(edge_file_path)}...")
            try:
                import pandas as pd
                edge_df = pd.read_parquet(edge_file_path)
                src = edge_df["source"].to_numpy(dtype=np.int64)
                dst = edge_df["target"].to_numpy(dtype=np.int64)
                w = edge_df["weight"].to_numpy(dtype=np.float32)  # int64 -> fp32 RN, reference :112
                self._build_from_edges(src, dst, w)
            except nat.NativeError:
                raise  # a missing CUDA library / device is not a data error: fail loudly
            except Exception as exc:  # noqa: BLE001 - reference :121-123 prints and falls back to empty
                print(f"    Error reading edge file {edge_file_path}: {exc}. Initializing empty graph.")
                self._initialize_empty_matrices()
        else:
            self._initialize_empty_matrices()

    # -- construction from in-memory edge arrays (what GraphBuilder uses; no parquet round trip)
    @classmethod
    def from_edge_arrays(cls, nodes: Dict[int, Any], src, dst, weight, epsilon_propagation: float = 1e-9,
                         n_value: Optional[int] = None, assume_coalesced: bool = False,
                         result_device="cpu") -> "DirectedNgramGraph":
        g = cls(nodes=nodes, edge_file_path=None, epsilon_propagation=epsilon_propagation, n_value=n_value)
        if g.number_of_nodes > 0 and len(src) > 0:
            g._build_from_edges(src, dst, weight, assume_coalesced=assume_coalesced, result_device=result_device)
        return g

    def _initialize_empty_matrices(self):
        self.number_of_edges = 0
        n = self.number_of_nodes
        self.A_out_w = _empty_coo(n)
        self.A_in_w = _empty_coo(n)
        self.A_undirected_norm_sparse = _empty_coo(n)
        self.mathcal_A_out = _empty_coo(n)
        self.mathcal_A_in = _empty_coo(n)

    def gcn_data(self, x: torch.Tensor, device, **extra):
        """Data object for ProtGramDirectGCN with the three propagation matrices attached the way
        the reference trainer does (protgram_directgcn_trainer.py:362-367).  When this graph was
        built in this process the device CSR from the normalisation kernels is handed to the layer
        directly, so the edge lists are not re-sorted."""
        from .protgram_directgcn import Data, register_symmetric_structure
        dev = torch.device(device)
        side = getattr(self, "_pg_device", None)
        if (side is None or side["pattern"].device != dev) and self.__dict__.get("_pg_sidecar") is not None:
            side = self._upload_sidecar(dev)      # graph loaded from disk with its CSR sidecar: no edge-list sort, 16 B / entry uploaded
        if side is not None and side["pattern"].device == dev:
            ei = side["pattern"]
            ew = (side["val_in"], side["val_out"], side["val_und"])
            register_symmetric_structure(ei, ew, self.number_of_nodes, side["rowptr"], side["col"], static=False)
        else:
            ei = self.mathcal_A_in.indices().to(dev)
            ew = tuple(t.values().to(dev) for t in (self.mathcal_A_in, self.mathcal_A_out, self.A_undirected_norm_sparse))
        return Data(x=x.to(dev), edge_index_in=ei, edge_weight_in=ew[0], edge_index_out=ei, edge_weight_out=ew[1],
                    edge_index_undirected_norm=ei, edge_weight_undirected_norm=ew[2], num_nodes=self.number_of_nodes, **extra)

    # -- on-disk hand-off (DataUtils.save_object / load_object)
    def propagation_csr_host(self):
        """The three propagation matrices as one shared-pattern CSR in host memory (numpy), or None when the matrices do not
        share one pattern (graphs not built by this class)."""
        n = self.number_of_nodes
        side = self.__dict__.get("_pg_device")
        if side is not None:
            return {"n": np.int64(n), "rowptr": side["rowptr"].cpu().numpy(), "col": side["col"].cpu().numpy(),
                    "val_in": side["val_in"].cpu().numpy(), "val_out": side["val_out"].cpu().numpy(),
                    "val_und": side["val_und"].cpu().numpy()}
        host = self.__dict__.get("_pg_sidecar")
        if host is not None:
            return dict(host)
        try:
            mats = [self.mathcal_A_in, self.mathcal_A_out, self.A_undirected_norm_sparse]
        except AttributeError:
            return None
        if n == 0 or any(m is None or m._nnz() == 0 for m in mats):
            return None
        idx = mats[0].indices()
        if any(m._nnz() != mats[0]._nnz() or not torch.equal(m.indices(), idx) for m in mats[1:]):
            return None
        idx = idx.cpu()
        rowptr = np.zeros(n + 1, dtype=np.int64)
        rowptr[1:] = np.cumsum(np.bincount(idx[0].numpy(), minlength=n))
        return {"n": np.int64(n), "rowptr": rowptr, "col": idx[1].numpy().astype(np.int32),
                "val_in": mats[0].values().cpu().numpy(), "val_out": mats[1].values().cpu().numpy(),
                "val_und": mats[2].values().cpu().numpy()}

    def attach_propagation_csr(self, csr: dict) -> None:
        """Accept a sidecar only if it describes THIS graph's pattern (node count, entry count, row lengths)."""
        n, p = self.number_of_nodes, int(self.mathcal_A_in._nnz())
        rowptr, col = np.asarray(csr["rowptr"]), np.asarray(csr["col"])
        ok = (int(csr["n"]) == n and rowptr.shape == (n + 1,) and rowptr.dtype == np.int64 and col.shape == (p,) and col.dtype == np.int32
              and int(rowptr[0]) == 0 and int(rowptr[-1]) == p
              and all(np.asarray(csr[k]).shape == (p,) and np.asarray(csr[k]).dtype == np.float32 for k in ("val_in", "val_out", "val_und")))
        if ok and p:
            idx = self.mathcal_A_in.indices()
            ok = bool(np.array_equal(np.diff(rowptr), np.bincount(idx[0].numpy(), minlength=n))) and \
                bool(np.array_equal(col[: 1 << 16], idx[1][: 1 << 16].numpy().astype(np.int32)))
        if not ok:
            raise ValueError("sidecar does not match the pickled graph")
        self.__dict__["_pg_sidecar"] = {k: np.ascontiguousarray(csr[k]) for k in ("n", "rowptr", "col", "val_in", "val_out", "val_und")}

    def _upload_sidecar(self, dev):
        host = self.__dict__["_pg_sidecar"]
        up = lambda k: torch.from_numpy(host[k]).to(dev, non_blocking=True)
        rowptr, col = up("rowptr"), up("col")
        side = {"rowptr": rowptr, "col": col, "val_in": up("val_in"), "val_out": up("val_out"), "val_und": up("val_und")}
        side["pattern"] = csr_to_coo_indices(rowptr, col, self.number_of_nodes)
        self.__dict__["_pg_device"] = side
        return side

    def _build_from_edges(self, src, dst, w, assume_coalesced: bool = False, result_device="cpu"):
        dev = nat.current_device()
        as_dev = lambda a, dt: (a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))).to(dev, dtype=dt)
        src_d, dst_d, w_d = as_dev(src, torch.int64), as_dev(dst, torch.int64), as_dev(w, torch.float32)
        n = self.number_of_nodes
        self.number_of_edges = int(src_d.numel())  # reference :116 (rows of the edge table)
        if not assume_coalesced:
            src_d, dst_d, w_d = device_coalesce(src_d, dst_d, w_d, n)
        if src_d.numel() == 0:
            self._initialize_empty_matrices()
            return
        res = device_normalize(src_d, dst_d, w_d, n, self.epsilon_propagation)
        pat = csr_to_coo_indices(res["rowptr"], res["col"], n)
        self._pg_device = {"pattern": pat, "rowptr": res["rowptr"], "col": res["col"], "val_in": res["val_in"],
                           "val_out": res["val_out"], "val_und": res["val_und"]}
        a_out_idx = torch.stack([src_d, dst_d])
        a_in_idx = torch.stack([res["in_src"], res["in_dst"]])
        tensors = [a_out_idx, w_d, a_in_idx, res["in_w"], pat, res["val_und"], res["val_out"], res["val_in"]]
        if torch.device(result_device).type == "cpu" and dev.type == "cuda":
            tensors, self.__dict__["_ready_event"] = _async_to_host(tensors)   # lands under whatever the caller enqueues next
        else:
            tensors = [t.to(result_device) for t in tensors]
        a_out_idx, w_o, a_in_idx, in_w, pat_o, v_und, v_out, v_in = tensors
        self.A_out_w = _coo(a_out_idx, w_o, n)
        self.A_in_w = _coo(a_in_idx, in_w, n)
        self.A_undirected_norm_sparse = _coo(pat_o, v_und, n)
        self.mathcal_A_out = _coo(pat_o, v_out, n)
        self.mathcal_A_in = _coo(pat_o, v_in, n)

    # reference :275-287 -- public: the trainer calls it after moving A_*_w to the device
    def _create_propagation_matrices_for_gcn(self):
        if self.number_of_nodes == 0:
            self._initialize_empty_matrices()
            return
        a = self.A_out_w
        target = a.device
        if a._nnz() == 0:
            self.mathcal_A_out = _empty_coo(self.number_of_nodes, target)
            self.mathcal_A_in = _empty_coo(self.number_of_nodes, target)
            return
        dev = target if target.type == "cuda" else nat.current_device()
        a = a.coalesce()
        idx, val = a.indices().to(dev), a.values().to(dev, dtype=torch.float32)
        res = device_normalize(idx[0].contiguous(), idx[1].contiguous(), val.contiguous(), self.number_of_nodes,
                               self.epsilon_propagation)
        pat = csr_to_coo_indices(res["rowptr"], res["col"], self.number_of_nodes).to(target)
        self.mathcal_A_out = _coo(pat, res["val_out"].to(target), self.number_of_nodes)
        self.mathcal_A_in = _coo(pat, res["val_in"].to(target), self.number_of_nodes)
