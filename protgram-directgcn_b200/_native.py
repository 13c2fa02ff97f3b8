"""ctypes binding of libpgb200.so (include/pgb200.h).  No fallback: if the library is missing or
there is no CUDA device, every hot-path call raises."""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_uint32, c_void_p
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpgb200.so")


class NativeError(RuntimeError):
    pass


PG_FASTA_ETOOSMALL, PG_FASTA_ENONASCII, PG_EPACK = -10, -11, -12   # include/pgb200.h


_P = c_void_p
_SIGS = {
    "pg_version": (c_int, []),
    "pg_last_error": (c_char_p, []),
    "pg_launch_count": (ctypes.c_uint64, []),
    "pg_byte_presence": (c_int, [_P, c_int64, _P, _P]),
    "pg_synth_corpus": (c_int, [_P, c_int64, c_int64, c_int, c_uint32, c_int, _P]),
    "pg_fasta_open": (_P, [ctypes.c_char_p]),
    "pg_fasta_close": (None, [_P]),
    "pg_fasta_file_bytes": (c_int64, [_P]),
    "pg_fasta_records": (c_int64, [_P]),
    "pg_fasta_stopped_early": (c_int, [_P]),
    "pg_fasta_next_chunk": (c_int64, [_P, _P, c_int64, c_int, c_int, c_int]),
    "pg_fasta_pack_parallel": (c_int64, [ctypes.c_char_p, _P, c_int64, c_int, c_int, c_int, c_int, _P, _P]),
    "pg_fasta_pack_window": (c_int64, [_P, _P, c_int64, c_int64, _P, c_int64, c_int, c_int, c_int, c_int, _P, _P]),
    "pg_pack5_bytes": (c_int64, [c_int64]),
    "pg_pack5_host": (c_int64, [_P, c_int64, _P]),
    "pg_unpack5": (c_int, [_P, c_int64, _P, _P]),
    "pg_ngram_count": (c_int, [_P, c_int64, c_int, _P, c_int, _P, _P, _P, c_size_t, _P]),
    "pg_ngram_count_ws_bytes": (c_size_t, [c_int, c_int]),
    "pg_ngram_count_ws_bytes_for": (c_size_t, [c_int, c_int, c_int64]),
    "pg_debug_count_variant": (None, [c_int]),
    "pg_graph_extract_ws_bytes": (c_size_t, [c_int, c_int]),
    "pg_graph_extract_sizes": (c_int, [_P, _P, c_int, c_int, _P, _P, c_size_t, _P]),
    "pg_graph_extract_fill": (c_int, [_P, c_int, c_int, c_int64, c_int64, _P, _P, _P, _P, _P, c_size_t, _P]),
    "pg_graph_extract_range_ws_bytes": (c_size_t, [c_int, c_int64]),
    "pg_graph_extract_range_mark": (c_int, [_P, c_int, c_int, c_int64, c_int64, _P, _P, _P, c_size_t, _P]),
    "pg_node_ids_ws_bytes": (c_size_t, [c_int64]),
    "pg_node_ids_from_presence": (c_int, [_P, c_int64, _P, _P, _P, c_size_t, _P]),
    "pg_node_codes_emit": (c_int, [_P, _P, c_int64, _P, _P]),
    "pg_graph_extract_range_fill": (c_int, [_P, c_int, c_int, c_int64, c_int64, _P, c_int64, _P, _P, _P, _P, c_size_t, _P]),
    "pg_sort_pairs_ws_bytes": (c_size_t, [c_int64]),
    "pg_sort_pairs": (c_int, [_P, _P, _P, _P, c_int64, c_int, _P, c_size_t, _P]),
    "pg_coo_coalesce_ws_bytes": (c_size_t, [c_int64]),
    "pg_coo_coalesce": (c_int, [_P, _P, _P, c_int64, c_int64, _P, _P, _P, _P, _P, c_size_t, _P]),
    "pg_normalize_ws_bytes": (c_size_t, [c_int64, c_int64]),
    "pg_normalize_sizes": (c_int, [_P, _P, _P, c_int64, c_int64, _P, _P, c_size_t, _P]),
    "pg_normalize_fill": (c_int, [_P, _P, _P, c_int64, c_int64, c_float, c_int64, _P, _P, _P, _P, _P, _P, _P, _P,
                                  _P, c_size_t, _P]),
    "pg_degree_sums_rows": (c_int, [_P, _P, c_int64, _P, _P, c_int64, c_int64, c_int64, _P, _P, _P]),
    "pg_normalize_rows_ws_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "pg_normalize_rows_sizes": (c_int, [_P, _P, c_int64, _P, _P, c_int64, c_int64, c_int64, c_int64, _P, _P, c_size_t, _P]),
    "pg_normalize_rows_structure": (c_int, [_P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, _P, _P, _P, _P, _P, _P,
                                            _P, c_size_t, _P]),
    "pg_normalize_rows_values": (c_int, [_P, _P, c_int64, c_int64, c_int64, c_int64, c_int64, _P, _P, _P, _P, c_float, _P, _P, _P,
                                         _P, c_size_t, _P]),
    "pg_rowptr_from_sorted": (c_int, [_P, c_int64, c_int64, _P, _P]),
    "pg_coo_from_csr": (c_int, [_P, _P, c_int64, c_int64, _P, _P, _P]),
    "pg_edges_to_csr_ws_bytes": (c_size_t, [c_int64]),
    "pg_edges_to_csr": (c_int, [_P, _P, _P, c_int64, c_int64, _P, _P, _P, _P, c_size_t, _P]),
    "pg_spmm_fanout": (c_int, [_P, _P, _P, _P, _P, c_int, c_int64, c_int, _P, c_int64, _P, c_int64, c_int64, _P, _P]),
    "pg_spmm_fanout_scaled": (c_int, [_P, _P, _P, _P, _P, c_int, c_int64, c_int, _P, c_int64, _P, c_int64, c_int64, _P, _P, _P, c_int,
                                      _P, _P]),
    "pg_spmm_fanout_split": (c_int, [_P, _P, _P, _P, _P, c_int, c_int64, c_int, _P, _P, c_int64, c_int64, c_int64, _P, _P, _P, c_int,
                                     _P, _P]),
    "pg_spmm_fanin_split": (c_int, [_P, _P, _P, _P, _P, c_int, c_int64, c_int, _P, c_int64, c_int64, _P, c_int64, _P, c_int64, c_int,
                                    _P, _P]),
    "pg_gather_rows": (c_int, [_P, c_int64, _P, c_int64, c_int, _P, c_int64, _P]),
    "pg_peer_alloc": (c_int, [c_size_t, _P, _P]),
    "pg_peer_open": (c_int, [_P, _P]),
    "pg_peer_close": (c_int, [_P]),
    "pg_peer_free": (c_int, [_P]),
    "pg_halo_push": (c_int, [_P, c_int64, _P, _P, _P, _P, c_int, c_int, c_int, c_int64, c_uint32, _P, _P]),
    "pg_halo_wait": (c_int, [_P, c_int, c_int, c_uint32, _P, _P]),
    "pg_spmm_fanin": (c_int, [_P, _P, _P, _P, _P, c_int, c_int64, c_int, _P, c_int64, c_int64, _P, c_int64, _P,
                              c_int64, c_int, _P, _P]),
    "pg_layer_gemm_fwd": (c_int, [_P, c_int64, _P, c_int64, _P, _P, _P, c_int, _P, _P, c_int64, c_int64, c_int,
                                  c_int, c_int, c_int, c_float, _P, c_int64, _P]),
    "pg_layer_gemm_fwd_tc_supported": (c_int, [c_int, c_int]),
    "pg_layer_gemm_fwd_tc_ws_bytes": (c_size_t, [c_int, c_int, c_int]),
    "pg_layer_gemm_fwd_tc": (c_int, [_P, c_int64, _P, c_int64, _P, _P, _P, c_int, _P, _P, c_int64, c_int64, c_int,
                                     c_int, c_int, c_int, c_float, _P, c_int64, _P, c_size_t, _P]),
    "pg_layer_gemm_fwd_tc_check": (c_int, [_P, c_int, c_int, c_int, _P]),
    "pg_lrelu_bwd": (c_int, [_P, _P, c_float, c_int64, _P, _P]),
    "pg_layer_gemm_bwd_data": (c_int, [_P, c_int64, _P, _P, c_int64, _P, _P, _P, c_int, c_int64, c_int, c_int,
                                       c_int, _P, c_int64, _P, c_int64, _P, _P]),
    "pg_layer_gemm_bwd_weight_ws_bytes": (c_size_t, [c_int64, c_int, c_int, c_int]),
    "pg_layer_gemm_bwd_weight": (c_int, [_P, c_int64, _P, c_int64, _P, _P, _P, c_int, _P, c_int64, c_int64, c_int,
                                         c_int, c_int, _P, _P, c_size_t, _P]),
    "pg_layer_gemm_bwd_data_tc_ws_bytes": (c_size_t, [c_int, c_int, c_int]),
    "pg_layer_gemm_bwd_data_tc": (c_int, [_P, c_int64, _P, _P, c_int64, _P, _P, _P, c_int, c_int64, c_int, c_int,
                                          c_int, _P, c_int64, _P, c_int64, _P, _P, c_size_t, _P]),
    "pg_layer_gemm_bwd_weight_tc_ws_bytes": (c_size_t, [c_int64, c_int, c_int, c_int]),
    "pg_layer_gemm_bwd_weight_tc": (c_int, [_P, c_int64, _P, c_int64, _P, _P, _P, c_int, _P, c_int64, c_int64, c_int,
                                            c_int, c_int, _P, _P, c_size_t, _P]),
    "pg_layer_gemm_bwd_dx_tc_ws_bytes": (c_size_t, [c_int, c_int, c_int]),
    "pg_layer_gemm_bwd_dx_tc": (c_int, [_P, c_int64, _P, c_int64, _P, c_int64, c_int, c_int, c_int, c_int, _P, c_int64, _P, c_size_t, _P]),
    "pg_layer_gate_grad_tc_ws_bytes": (c_size_t, [c_int64, c_int, c_int]),
    "pg_layer_gate_grad_tc": (c_int, [_P, c_int64, _P, _P, c_int64, c_int64, c_int, c_int, c_int, _P, _P, c_size_t, _P]),
    "pg_tc_check": (c_int, [_P, c_size_t, _P]),
    "pg_layer_gate_grad_ws_bytes": (c_size_t, [c_int64, c_int, c_int]),
    "pg_layer_gate_grad": (c_int, [_P, c_int64, _P, _P, c_int64, c_int64, c_int, c_int, c_int, _P, _P, c_size_t, _P]),
    "pg_layer_gemm_bwd_dx": (c_int, [_P, c_int64, _P, c_int64, _P, c_int64, c_int, c_int, c_int, c_int, _P, c_int64, _P]),
    "pg_decoder_grads_supported": (c_int, [c_int]),
    "pg_decoder_grads_ws_bytes": (c_size_t, [c_int64, c_int, c_int]),
    "pg_decoder_grads": (c_int, [_P, c_int64, _P, c_int64, _P, c_int64, c_int, c_int, c_float, _P, _P, _P, c_size_t, _P]),
    "pg_linear_fwd": (c_int, [_P, c_int64, c_int64, c_int, _P, _P, c_int, c_int, _P, c_int64, _P]),
    "pg_linear_bwd_data": (c_int, [_P, c_int64, c_int64, c_int, _P, c_int, _P, c_int64, _P]),
    "pg_linear_bwd_weight_ws_bytes": (c_size_t, [c_int64, c_int, c_int]),
    "pg_linear_bwd_weight": (c_int, [_P, c_int64, _P, c_int64, c_int64, c_int, c_int, _P, _P, c_size_t, _P]),
    "pg_colsum_ws_bytes": (c_size_t, [c_int64, c_int]),
    "pg_colsum": (c_int, [_P, c_int64, c_int64, c_int, _P, _P, c_size_t, _P]),
    "pg_linear_tc_ws_bytes": (c_size_t, [c_int, c_int]),
    "pg_linear_tc": (c_int, [_P, c_int64, c_int64, c_int, _P, _P, c_int, _P, c_int64, _P, c_size_t, _P]),
    "pg_l2_normalize_rows": (c_int, [_P, c_int64, c_int64, c_int, c_float, _P, c_int64, _P]),
    "pg_pack_layer_params": (c_int, [_P, c_int64, c_int, c_int, c_int, _P, _P, _P, _P, _P]),
    "pg_unpack_layer_param_grads": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, _P, _P]),
    "pg_next_node_labels": (c_int, [_P, _P, _P, c_int64, _P, _P]),
    "pg_ngram_feature_init": (c_int, [_P, c_int64, _P, c_int64, c_int, c_int, _P, c_int64, c_int, _P, c_int64, _P]),
    "pg_subgraph_ws_bytes": (c_size_t, [c_int64]),
    "pg_subgraph_sizes": (c_int, [_P, _P, c_int64, _P, c_int64, _P, _P, _P, c_size_t, _P]),
    "pg_subgraph_fill": (c_int, [_P, _P, _P, _P, _P, c_int64, _P, c_int64, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "pg_pool_proteins": (c_int, [_P, _P, c_int64, c_int, _P, c_int, _P, _P, c_int64, c_int, _P, c_int64, _P, _P]),
    "pg_softmax_nll_ws_bytes": (c_size_t, [c_int64, c_int64]),
    "pg_softmax_nll": (c_int, [_P, c_int64, c_int64, c_int64, _P, c_float, _P, _P, _P, _P, c_size_t, _P]),
}

LAYER_PARAM_FIELDS = ("w_in", "w_out", "w_und", "w_sh", "b_in", "b_out", "b_und", "bs_in", "bs_out", "bs_und", "w_res", "b_res",
                      "c_in", "c_out", "c_dir", "c_und", "c_all")


class LayerParamsStruct(ctypes.Structure):
    """Mirror of `pg_layer_params` / `pg_layer_param_grads` (include/pgb200.h): 17 device pointers."""
    _fields_ = [(name, _P) for name in LAYER_PARAM_FIELDS]


def layer_params(**tensors):
    """-> (by-reference argument, keep-alive) for pg_pack_layer_params / pg_unpack_layer_param_grads."""
    st = LayerParamsStruct(**{k: ptr(tensors.get(k)) for k in LAYER_PARAM_FIELDS})
    return ctypes.byref(st), (st, tensors)


class SpmmOperandStruct(ctypes.Structure):
    """Mirror of `pg_spmm_operand` (include/pgb200.h)."""
    _fields_ = [("lo", _P), ("ld_lo", c_int64), ("hi", _P), ("ld_hi", c_int64), ("split", c_int64)]


def spmm_operand(lo: torch.Tensor, hi: Optional[torch.Tensor], split: int):
    """-> by-reference `pg_spmm_operand`: rows < split from `lo`, the others from `hi` (2-D fp32 views; their strides and
    storage offsets are honoured, so a column chunk is just `t[:, c0:c0 + w]`)."""
    st = SpmmOperandStruct(ptr(lo), lo.stride(0), ptr(hi) if hi is not None and hi.numel() else None,
                           hi.stride(0) if hi is not None and hi.numel() else 0, int(split))
    st._keep = (lo, hi)
    return ctypes.byref(st)


class DeviceView:
    """A device buffer that torch did not allocate (a CUDA-IPC receive slot), exposed through `__cuda_array_interface__` so that
    `torch.as_tensor(view, device=...)` wraps it without a copy."""

    def __init__(self, pointer: int, shape, typestr: str = "<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(int(v) for v in shape), "typestr": typestr, "data": (int(pointer), False),
                                         "version": 2, "strides": None}


def view_tensor(pointer: int, shape, device) -> torch.Tensor:
    """fp32 tensor over raw device memory at `pointer` (kept alive by its owner, not by the tensor)."""
    if int(shape[0]) == 0 or pointer == 0:
        return torch.empty(tuple(shape), dtype=torch.float32, device=device)
    return torch.as_tensor(DeviceView(pointer, shape), device=device)


def peer_alloc(nbytes: int):
    """-> (device pointer, 64-byte IPC handle as bytes) of a zeroed cudaMalloc'ed buffer peers can map (pg_peer_alloc)."""
    ptr_out = c_void_p()
    handle = (ctypes.c_ubyte * 64)()
    call("pg_peer_alloc", c_size_t(max(256, int(nbytes))), ctypes.byref(ptr_out), handle)
    return int(ptr_out.value), bytes(handle)


def peer_open(handle: bytes) -> int:
    ptr_out = c_void_p()
    buf = (ctypes.c_ubyte * 64).from_buffer_copy(handle)
    call("pg_peer_open", buf, ctypes.byref(ptr_out))
    return int(ptr_out.value)


class SpmmPlanStruct(ctypes.Structure):
    """Mirror of `pg_spmm_plan` (include/pgb200.h)."""
    _fields_ = [("chunk", ctypes.c_int32), ("n_long", c_int64), ("n_items", c_int64), ("d_long_rows", _P),
                ("d_item_ptr", _P), ("d_item_row", _P), ("d_partials", _P)]


class SpmmPlan:
    """Long-row work split for the SpMM kernels, built once per CSR structure (depends on rowptr only).
    Keeps its device arrays alive; `.ref(width)` returns the pointer to pass as `plan` (None if the
    structure has no long rows)."""

    CHUNK = int(os.environ.get("PGB200_SPMM_CHUNK", "512"))

    def __init__(self, rowptr: torch.Tensor, chunk: Optional[int] = None):
        chunk = self.CHUNK if chunk is None else chunk
        deg = rowptr[1:] - rowptr[:-1]
        long_rows = torch.nonzero(deg > chunk).flatten()
        self.chunk = int(chunk)
        self.n_long = int(long_rows.numel())
        self.n_items = 0
        self._struct = None
        self._partials = None
        if self.n_long:
            per = (deg[long_rows] + chunk - 1) // chunk
            item_ptr = torch.zeros(self.n_long + 1, dtype=torch.int64, device=rowptr.device)
            item_ptr[1:] = torch.cumsum(per, 0)
            self.n_items = int(item_ptr[-1])
            self.long_rows = long_rows.to(torch.int32).contiguous()
            self.item_ptr = item_ptr
            self.item_row = torch.repeat_interleave(torch.arange(self.n_long, device=rowptr.device, dtype=torch.int32), per).contiguous()

    def ref(self, width: int):
        if not self.n_long:
            return None
        need = self.n_items * int(width)
        if self._partials is None or self._partials.numel() < need:
            self._partials = torch.empty(need, dtype=torch.float32, device=self.long_rows.device)
            self._struct = None
        if self._struct is None:
            self._struct = SpmmPlanStruct(self.chunk, self.n_long, self.n_items, ptr(self.long_rows), ptr(self.item_ptr),
                                          ptr(self.item_row), ptr(self._partials))
        return ctypes.byref(self._struct)


SOFTMAX_NLL_MAX_CLASSES = 28672  # PG_SOFTMAX_NLL_MAX_CLASSES (include/pgb200.h)

_lib = None
launches = 0  # number of C-ABI calls that enqueue GPU work (bench.py reports it)


def exported_symbols():
    return sorted(_SIGS)


def load():
    """Load the shared library (does not need a GPU).  Raises NativeError if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                f"{LIB_PATH} is missing: build it with `python protgram-directgcn_b200/build.py` "
                "(there is no CPU fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def require_cuda():
    if not torch.cuda.is_available():
        raise NativeError("protgram-directgcn_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def current_device() -> torch.device:
    require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def check_tensor(t: torch.Tensor, what: str = "input"):
    """Hot-path inputs must live on the GPU: there is no CPU implementation to fall back to."""
    if not t.is_cuda:
        raise NativeError(f"{what} must be a CUDA tensor: the CUDA path has no CPU fallback")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t: Optional[torch.Tensor]):
    if t is None:
        return None
    if not t.is_cuda:
        raise NativeError("expected a CUDA tensor")
    return t.data_ptr()


def call(name: str, *args):
    """Invoke an int-returning entry point; raise NativeError with pg_last_error() on failure."""
    global launches
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise NativeError(f"{name} failed ({rc}): {lib.pg_last_error().decode()}")
    launches += 1
    return rc


def query(name: str, *args) -> int:
    return int(getattr(load(), name)(*args))


_CONSTS = {"PG_MAX_PEERS": 16}     # include/pgb200.h


def query_const(name: str) -> int:
    return _CONSTS[name]


def kernel_launches() -> int:
    """Kernels launched by libpgb200.so in this process (bench.py's gpu_launches)."""
    return int(load().pg_launch_count())


def workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
