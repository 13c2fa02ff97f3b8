"""DirectGCNLayer / ProtGramDirectGCN -- drop-in for the reference's
src/models/protgram_directgcn.py (same constructor signatures, parameter names => identical
state_dict keys, same forward contracts), with the propagation and the dense transform running
through libpgb200.so.

Reference layer (:93-135): 6 x (gather -> scale -> scatter_add) + 4 Linears + 6 bias adds + 5 gate
multiplies + constant.  Here (algebra checked to 7.8e-7):

    Z      = [A_in X | A_out X | U X]                       one fan-out SpMM (X gathered once)
    Y      = [a Z_in | b Z_out | c Z_und | X | a b c | 1] @ W_ext + constant     one fused GEMM
    a = C_all*C_directed*C_in, b = C_all*C_directed*C_out, c = C_all*C_undirected
    W_ext  = [ (W_in+W_sh)^T ; (W_out+W_sh)^T ; (W_und+W_sh)^T ; W_res^T ; beta_in ; beta_out ; beta_und ; b_res ]

ProtGramDirectGCN additionally folds res_proj, the residual add and leaky_relu (:210-215) into
the same GEMM epilogue.  Backward = lrelu' -> dW_ext GEMM (split over rows) -> dA GEMM + gate
dot-products -> fan-in SpMM over the source-grouped structure (csrc/{spmm,gemm}.cu).
No CPU fallback: CPU tensors raise.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _native as nat
from .models_utils import EmbeddingProcessor


class Data:
    """Minimal stand-in for torch_geometric.data.Data (attribute bag with .to()); the model only
    uses getattr() on it (reference :196-203), so a real PyG Data works as well."""

    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    def to(self, device):
        for k, v in list(self.__dict__.items()):
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self


# ------------------------------------------------------------------------------------------------
# edge lists -> cached CSR structures
# ------------------------------------------------------------------------------------------------
class _Csr:
    __slots__ = ("rowptr", "col", "vals", "_plan")

    def __init__(self, rowptr, col, vals):
        self.rowptr, self.col, self.vals = rowptr, col, vals
        self._plan = None

    def plan(self, width: int):
        """Long-row split for skewed degree distributions (built on first use, cached)."""
        if self._plan is None:
            self._plan = nat.SpmmPlan(self.rowptr)
        return self._plan.ref(width)


def _edges_to_csr(group: torch.Tensor, other: torch.Tensor, w: Optional[torch.Tensor], n: int) -> _Csr:
    nnz = group.numel()
    dev = group.device
    rowptr = torch.empty(n + 1, dtype=torch.int64, device=dev)
    col = torch.empty(nnz, dtype=torch.int32, device=dev)
    val = torch.empty(nnz, dtype=torch.float32, device=dev)
    ws = nat.workspace(nat.query("pg_edges_to_csr_ws_bytes", nnz), dev)
    nat.call("pg_edges_to_csr", nat.ptr(group), nat.ptr(other), nat.ptr(w), nnz, n, nat.ptr(rowptr), nat.ptr(col), nat.ptr(val),
             nat.ptr(ws), ws.numel(), nat.stream_ptr())
    return _Csr(rowptr, col, [val])


class EdgeStructure:
    """CSR views of the three edge lists of one graph.

    by_dst: rows = message targets (ei[1]), cols = sources (ei[0])  -> forward  out[t] += w*x[s]
    by_src: rows = sources, cols = targets                          -> backward dx[s] += w*dy[t]
    `shared` = the three edge_index tensors are identical (reference-built graphs): one pattern,
    three value arrays, so the nv=3 kernels gather each neighbour row once."""

    def __init__(self, eis, ews, n: int):
        self.n = n
        same = all(e.shape == eis[0].shape for e in eis) and all(
            e.data_ptr() == eis[0].data_ptr() or torch.equal(e, eis[0]) for e in eis[1:])
        self.shared = bool(same)
        prep = lambda e: e.to(torch.int64).contiguous()
        fw = lambda w: None if w is None else w.to(torch.float32).contiguous()
        self.by_dst: List[_Csr] = []
        self.by_src: List[_Csr] = []
        for ei, ew in zip(eis, ews):
            ei = prep(ei)
            self.by_dst.append(_edges_to_csr(ei[1], ei[0], fw(ew), n))
            self.by_src.append(_edges_to_csr(ei[0], ei[1], fw(ew), n))
        if self.shared:
            # stable sort => the permutation is the same for the three value arrays
            for lst in (self.by_dst, self.by_src):
                merged = _Csr(lst[0].rowptr, lst[0].col, [c.vals[0] for c in lst])
                lst[:] = [merged]
            d, s = self.by_dst[0], self.by_src[0]
            if torch.equal(d.rowptr, s.rowptr) and torch.equal(d.col, s.col) and all(
                    torch.equal(a, b) for a, b in zip(d.vals, s.vals)):
                self.by_src = self.by_dst  # symmetric matrices: transposed structure == structure
        self.nnz_total = sum(int(e.shape[1]) for e in eis)


import collections as _collections

_STRUCT_CACHE: "_collections.OrderedDict[tuple, EdgeStructure]" = _collections.OrderedDict()
STRUCT_CACHE_ENTRIES = 64      # least-recently-used eviction; pre-registered structures also ride on their edge tensor (tag)


def _cache_key(eis, ews, n):
    return tuple((t.data_ptr(), tuple(t.shape)) if t is not None else None for t in (*eis, *ews)) + (n,)


def _versions(eis, ews):
    return tuple(t._version if t is not None else -1 for t in (*eis, *ews))


def _cache_put(key, st) -> None:
    _STRUCT_CACHE[key] = st
    _STRUCT_CACHE.move_to_end(key)
    while len(_STRUCT_CACHE) > STRUCT_CACHE_ENTRIES:
        _STRUCT_CACHE.popitem(last=False)


def _ews_key(ews):
    return tuple((t.data_ptr(), tuple(t.shape)) if t is not None else None for t in ews)


def register_symmetric_structure(ei: torch.Tensor, ews, n: int, rowptr: torch.Tensor, col: torch.Tensor,
                                 static: bool = True) -> None:
    """Hand the layer a ready CSR for a shared, row-major sorted, value-symmetric pattern (what the
    normalisation kernels emit for reference-built graphs): ews = (in, out, undirected) values in
    pattern order.  Symmetry makes grouped-by-target == grouped-by-source == the row CSR.
    static=True: the buffers are rewritten in place by their owner (CUDA-graph replay), so the
    cached structure stays valid across version bumps of the tensors.
    The structure is attached to `ei` itself (not to the LRU cache of structures built from edge lists): it
    lives exactly as long as the edge tensor the caller hands to the model, whatever passes through that cache."""
    st = EdgeStructure.__new__(EdgeStructure)
    st.n, st.shared = n, True
    csr = _Csr(rowptr, col, [w.contiguous() for w in ews])
    st.by_dst, st.by_src = [csr], [csr]
    st.nnz_total = 3 * int(col.numel())
    st._keepalive = (ews,)          # not `ei`: the tag below would make that a reference cycle
    st._static = static
    st._vers = _versions((ei, ei, ei), ews)
    st._tag_ews = _ews_key(ews)
    st._tag_n = n
    ei._pg_struct = st


def register_structure(eis, ews, n: int, struct) -> None:
    """Attach a ready structure object (e.g. partitioned.PartitionedStructure) to the edge tensors a Data object carries:
    the layer looks at the tag before it consults the cache of CSRs it builds itself."""
    struct._tag_ews = _ews_key(ews)
    struct._tag_n = n
    eis[0]._pg_struct = struct


def get_structure(eis, ews, n: int) -> EdgeStructure:
    tagged = getattr(eis[0], "_pg_struct", None)    # set by register_*structure: survives cache eviction
    if tagged is not None and all(e is eis[0] for e in eis[1:]) and getattr(tagged, "_tag_n", n) == n \
            and getattr(tagged, "_tag_ews", None) in (None, _ews_key(ews)):
        return tagged
    key = _cache_key(eis, ews, n)
    st = _STRUCT_CACHE.get(key)
    if st is not None and not getattr(st, "_static", False) and st._vers != _versions(eis, ews):
        st = None  # tensors were modified in place since the CSR was built
    if st is None:
        if eis[0].numel() == 2 and eis[0].shape[-1] == 1 and getattr(eis[0], "_pg_placeholder", False):
            raise RuntimeError("placeholder edge_index without its pre-registered structure (it must be registered with "
                               "register_symmetric_structure before the model sees it)")
        st = EdgeStructure(eis, ews, n)
        st._keepalive = (eis, ews)  # data_ptr keys stay valid while cached
        st._static = False
        st._vers = _versions(eis, ews)
    _cache_put(key, st)
    return st


# ------------------------------------------------------------------------------------------------
# kernels as autograd function
# ------------------------------------------------------------------------------------------------
def _fanout(struct: EdgeStructure, x: torch.Tensor, f_in: int, scales=None, scale_stride: int = 1, transposed: bool = False) -> torch.Tensor:
    """Z = [A_in X | A_out X | U X]; with `scales` = (s_in, s_out, s_und) indexed by the GATHERED row: A_v diag(s_v) X.
    transposed=True runs over the source-grouped structure: Z_v = A_v^T diag(s_v) X (the regrouped input gradient of the
    backward pass; for value-symmetric matrices both structures are the same object)."""
    if getattr(struct, "partitioned", False):   # rows of this rank only; neighbour rows arrive by exchange (host/partitioned.py)
        return struct.fanout(x, f_in, scales, scale_stride)
    n = x.shape[0]
    z = torch.empty((n, 3 * f_in), dtype=torch.float32, device=x.device)
    st = nat.stream_ptr()
    csrs = struct.by_src if transposed else struct.by_dst
    if struct.shared and scales is not None:
        c = csrs[0]
        nat.call("pg_spmm_fanout_scaled", nat.ptr(c.rowptr), nat.ptr(c.col), nat.ptr(c.vals[0]), nat.ptr(c.vals[1]), nat.ptr(c.vals[2]),
                 3, n, f_in, nat.ptr(x), x.stride(0), nat.ptr(z), z.stride(0), 0, nat.ptr(scales[0]), nat.ptr(scales[1]),
                 nat.ptr(scales[2]), int(scale_stride), c.plan(3 * f_in), st)
    elif struct.shared:
        c = csrs[0]
        nat.call("pg_spmm_fanout", nat.ptr(c.rowptr), nat.ptr(c.col), nat.ptr(c.vals[0]), nat.ptr(c.vals[1]), nat.ptr(c.vals[2]),
                 3, n, f_in, nat.ptr(x), x.stride(0), nat.ptr(z), z.stride(0), 0, c.plan(3 * f_in), st)
    else:
        for v, c in enumerate(csrs):
            if scales is not None:
                nat.call("pg_spmm_fanout_scaled", nat.ptr(c.rowptr), nat.ptr(c.col), nat.ptr(c.vals[0]), None, None, 1, n, f_in,
                         nat.ptr(x), x.stride(0), nat.ptr(z), z.stride(0), v * f_in, nat.ptr(scales[v]), None, None, int(scale_stride),
                         c.plan(3 * f_in), st)
            else:
                nat.call("pg_spmm_fanout", nat.ptr(c.rowptr), nat.ptr(c.col), nat.ptr(c.vals[0]), None, None, 1, n, f_in,
                         nat.ptr(x), x.stride(0), nat.ptr(z), z.stride(0), v * f_in, c.plan(3 * f_in), st)
    return z


def _fanin(struct: EdgeStructure, dz: torch.Tensor, f_in: int, init: Optional[torch.Tensor]) -> torch.Tensor:
    if getattr(struct, "partitioned", False):
        return struct.fanin(dz, f_in, init)
    n = dz.shape[0]
    dx = torch.empty((n, f_in), dtype=torch.float32, device=dz.device)
    st = nat.stream_ptr()
    ldinit = init.stride(0) if init is not None else 0
    if struct.shared:
        c = struct.by_src[0]
        nat.call("pg_spmm_fanin", nat.ptr(c.rowptr), nat.ptr(c.col), nat.ptr(c.vals[0]), nat.ptr(c.vals[1]), nat.ptr(c.vals[2]),
                 3, n, f_in, nat.ptr(dz), dz.stride(0), 0, nat.ptr(init), ldinit, nat.ptr(dx), dx.stride(0), 0,
                 c.plan(3 * f_in), st)
    else:
        for v, c in enumerate(struct.by_src):
            nat.call("pg_spmm_fanin", nat.ptr(c.rowptr), nat.ptr(c.col), nat.ptr(c.vals[0]), None, None, 1, n, f_in,
                     nat.ptr(dz), dz.stride(0), v * f_in, nat.ptr(init) if v == 0 else None, ldinit, nat.ptr(dx),
                     dx.stride(0), 0 if v == 0 else 1, c.plan(3 * f_in), st)
    return dx


# Dense transform backend: "auto" = tcgen05 tensor cores (3 x TF32 split, fp32-level accuracy) when the
# hidden width makes it a real contraction (F_out >= 128, north_star) and the shape is supported,
# SIMT fp32 otherwise; "off" / "force" for tests and A/B timing.
import os as _os

TC_MODE = _os.environ.get("PGB200_TC", "auto")
TC_MIN_WIDTH, TC_MIN_ROWS = 128, 4096
BWD_DX_MODE = _os.environ.get("PGB200_BWD_DX", "fanout")   # "fanin": gather the gated gradient (the SIMT path's way) also on the TC path
DECODER_GRADS = _os.environ.get("PGB200_DECODER_GRADS", "auto")    # "auto" | "fused" | "simt" | "library" (see _LinearLogSoftmaxNLL.backward)
DECODER_GRADS_FUSED_MAX_CLASSES = 1024
TC_BWD_WEIGHT_MIN_ROWS = 65536   # the weight gradient splits the ROWS over CTAs: below this the SIMT kernel (more, smaller tiles) is faster


def _use_tensor_cores(n: int, f_in: int, f_out: int) -> bool:
    if TC_MODE == "off" or not nat.query("pg_layer_gemm_fwd_tc_supported", f_in, f_out):
        return False
    return TC_MODE == "force" or (f_out >= TC_MIN_WIDTH and n >= TC_MIN_ROWS)


def _finish_backward(ctx, ga, dgate, dx, dw, dy):
    """Common tail of _DirectGCNFused.backward: per-row gate gradients -> the gates' shapes, constant's gradient = dY."""
    if ctx.gate_stride == 1:
        dga, dgb, dgc = (dgate[v].reshape(ga.shape) for v in range(3))
    else:
        dga, dgb, dgc = (dgate[v].sum().reshape(ga.shape) for v in range(3))
    dconst = dy if ctx.has_const else None
    return dx, dga, dgb, dgc, dw, dconst, None, None, None, None


class _DirectGCNFused(torch.autograd.Function):
    """H = act( [aZ_in | bZ_out | cZ_und | X? | a b c | 1?] @ W_ext (+X) + constant )."""

    @staticmethod
    def forward(ctx, x, ga, gb, gc, w_ext, const_rows, struct, has_res, add_identity, slope):
        nat.check_tensor(x, "x")
        x = x.contiguous().float()
        ga, gb, gc = (g.contiguous().float() for g in (ga, gb, gc))
        w_ext = w_ext.contiguous().float()
        n, f_in = x.shape
        f_out = w_ext.shape[1]
        gate_stride = 1 if ga.numel() == n else 0
        if ga.numel() not in (1, n):
            raise ValueError(f"gate vectors must have 1 or {n} entries, got {ga.numel()}")
        z = _fanout(struct, x, f_in)
        h = torch.empty((n, f_out), dtype=torch.float32, device=x.device)
        if const_rows is not None:
            const_rows = const_rows.contiguous().float()
        ldc = const_rows.stride(0) if const_rows is not None else 0
        use_tc = _use_tensor_cores(n, f_in, f_out) and ldc % 4 == 0
        if use_tc:
            ws = nat.workspace(nat.query("pg_layer_gemm_fwd_tc_ws_bytes", f_in, f_out, int(has_res)), x.device)
            nat.call("pg_layer_gemm_fwd_tc", nat.ptr(z), z.stride(0), nat.ptr(x), x.stride(0), nat.ptr(ga), nat.ptr(gb), nat.ptr(gc),
                     gate_stride, nat.ptr(w_ext), nat.ptr(const_rows), ldc, n, f_in, f_out, int(has_res), int(add_identity),
                     float(slope), nat.ptr(h), h.stride(0), nat.ptr(ws), ws.numel(), nat.stream_ptr())
        else:
            nat.call("pg_layer_gemm_fwd", nat.ptr(z), z.stride(0), nat.ptr(x), x.stride(0), nat.ptr(ga), nat.ptr(gb), nat.ptr(gc),
                     gate_stride, nat.ptr(w_ext), nat.ptr(const_rows), ldc, n, f_in, f_out, int(has_res), int(add_identity),
                     float(slope), nat.ptr(h), h.stride(0), nat.stream_ptr())
        ctx.save_for_backward(x, ga, gb, gc, w_ext, z, h)
        ctx.use_tc = use_tc
        # the regrouped backward needs 128-bit rows on the tensor-core path; the SIMT kernels take any width
        ctx.bwd_regrouped = (f_in % 4 == 0 and f_out % 4 == 0) if use_tc else True
        ctx.struct, ctx.has_res, ctx.add_identity, ctx.slope = struct, bool(has_res), bool(add_identity), float(slope)
        ctx.gate_stride, ctx.has_const = gate_stride, const_rows is not None
        return h

    @staticmethod
    def backward(ctx, dh):
        x, ga, gb, gc, w_ext, z, h = ctx.saved_tensors
        n, f_in = x.shape
        f_out = w_ext.shape[1]
        st = nat.stream_ptr()
        dh = dh.contiguous().float()
        if ctx.slope != 1.0:
            dy = torch.empty_like(dh)
            nat.call("pg_lrelu_bwd", nat.ptr(dh), nat.ptr(h), ctx.slope, dh.numel(), nat.ptr(dy), st)
        else:
            dy = dh
        has_res = int(ctx.has_res)
        fanout_bwd = ctx.bwd_regrouped and BWD_DX_MODE == "fanout"
        # row-partitioned graphs: post the halo exchange of dY (and of the gates) NOW; the two GEMMs below run under it
        pending = None
        if fanout_bwd and ctx.needs_input_grad[0] and getattr(ctx.struct, "partitioned", False) and hasattr(ctx.struct, "fanout_begin"):
            pending = ctx.struct.fanout_begin(dy, f_out, scales=(ga, gb, gc), scale_stride=ctx.gate_stride)
        # dW_ext = A_ext^T dY
        dw = torch.empty_like(w_ext)
        tc = "_tc" if ctx.use_tc and (n >= TC_BWD_WEIGHT_MIN_ROWS or TC_MODE == "force") else ""
        ws = nat.workspace(nat.query(f"pg_layer_gemm_bwd_weight{tc}_ws_bytes", n, f_in, f_out, has_res), x.device)
        nat.call(f"pg_layer_gemm_bwd_weight{tc}", nat.ptr(z), z.stride(0), nat.ptr(x), x.stride(0), nat.ptr(ga), nat.ptr(gb),
                 nat.ptr(gc), ctx.gate_stride, nat.ptr(dy), dy.stride(0), n, f_in, f_out, has_res, nat.ptr(dw), nat.ptr(ws),
                 ws.numel(), st)
        dgate = torch.empty((3, n), dtype=torch.float32, device=x.device)
        if fanout_bwd:
            # gate gradients from the data-gradient GEMM with a dot-product epilogue (dZ is never written); the input gradient
            # regrouped as dX = sum_v (A_v^T (g_v * dY)) W'_v^T (+ residual): the gather runs over the SOURCE-grouped structure
            # (== the forward structure for value-symmetric matrices) and moves F_out-wide rows of dY once for all three matrices
            # instead of the fan-in kernel gathering the 3 F_in-wide gated gradient.  Tensor-core and SIMT GEMMs alike.
            t_ = "_tc" if ctx.use_tc else ""
            wsg = nat.workspace(nat.query(f"pg_layer_gate_grad{t_}_ws_bytes", n, f_in, f_out), x.device)
            nat.call(f"pg_layer_gate_grad{t_}", nat.ptr(dy), dy.stride(0), nat.ptr(w_ext), nat.ptr(z), z.stride(0), n, f_in, f_out, has_res,
                     nat.ptr(dgate), nat.ptr(wsg), wsg.numel(), st)
            dx = None
            if ctx.needs_input_grad[0]:
                if pending is not None:
                    t = ctx.struct.fanout_finish(pending)
                else:
                    t = _fanout(ctx.struct, dy, f_out, scales=(ga, gb, gc), scale_stride=ctx.gate_stride, transposed=True)
                dx = torch.empty((n, f_in), dtype=torch.float32, device=x.device)
                if ctx.use_tc:
                    ws3 = nat.workspace(nat.query("pg_layer_gemm_bwd_dx_tc_ws_bytes", f_in, f_out, has_res), x.device)
                    nat.call("pg_layer_gemm_bwd_dx_tc", nat.ptr(t), t.stride(0), nat.ptr(dy), dy.stride(0), nat.ptr(w_ext), n, f_in, f_out,
                             has_res, int(ctx.add_identity), nat.ptr(dx), dx.stride(0), nat.ptr(ws3), ws3.numel(), st)
                else:
                    nat.call("pg_layer_gemm_bwd_dx", nat.ptr(t), t.stride(0), nat.ptr(dy), dy.stride(0), nat.ptr(w_ext), n, f_in, f_out,
                             has_res, int(ctx.add_identity), nat.ptr(dx), dx.stride(0), st)
            return _finish_backward(ctx, ga, dgate, dx, dw, dy)
        # dA = dY W_ext^T  -> dZ (gated), dXres, dgates
        dz = torch.empty_like(z)
        dxres = torch.empty_like(x) if ctx.has_res else None
        args = (nat.ptr(dy), dy.stride(0), nat.ptr(w_ext), nat.ptr(z), z.stride(0), nat.ptr(ga), nat.ptr(gb), nat.ptr(gc),
                ctx.gate_stride, n, f_in, f_out, has_res, nat.ptr(dz), dz.stride(0), nat.ptr(dxres),
                dxres.stride(0) if dxres is not None else 0, nat.ptr(dgate))
        if ctx.use_tc:
            ws2 = nat.workspace(nat.query("pg_layer_gemm_bwd_data_tc_ws_bytes", f_in, f_out, has_res), x.device)
            nat.call("pg_layer_gemm_bwd_data_tc", *args, nat.ptr(ws2), ws2.numel(), st)
        else:
            nat.call("pg_layer_gemm_bwd_data", *args, st)
        dx = None
        if ctx.needs_input_grad[0]:
            init = dxres if ctx.has_res else (dy if ctx.add_identity else None)
            dx = _fanin(ctx.struct, dz, f_in, init)
        return _finish_backward(ctx, ga, dgate, dx, dw, dy)


# ------------------------------------------------------------------------------------------------
# parameter packing: 17 reference parameters -> (W_ext, a, b, c), one kernel each way
# ------------------------------------------------------------------------------------------------
class _PackLayerParams(torch.autograd.Function):
    """(W_ext, a, b, c) of one layer from the reference's parameters (csrc/params.cu).  Replaces ~12 tiny
    tensor ops in forward and ~25 in autograd's backward per layer; bit-identical in forward."""

    @staticmethod
    def forward(ctx, has_res, *params):
        names = nat.LAYER_PARAM_FIELDS
        t = {k: (None if v is None else v.detach().contiguous().float()) for k, v in zip(names, params)}
        w_in = t["w_in"]
        f_out, f_in = w_in.shape
        num_gate = t["c_in"].numel()
        k_ext = 3 * f_in + (f_in + 1 if has_res else 0) + 3
        dev = w_in.device
        w_ext = torch.empty((k_ext, f_out), dtype=torch.float32, device=dev)
        ga, gb, gc = (torch.empty(num_gate, dtype=torch.float32, device=dev) for _ in range(3))
        ref, keep = nat.layer_params(**t)
        nat.call("pg_pack_layer_params", ref, num_gate, f_in, f_out, int(has_res), nat.ptr(w_ext), nat.ptr(ga), nat.ptr(gb),
                 nat.ptr(gc), nat.stream_ptr())
        ctx.tensors, ctx.has_res, ctx.shapes = t, bool(has_res), [None if v is None else v.shape for v in params]
        return w_ext, ga, gb, gc

    @staticmethod
    def backward(ctx, dw_ext, dga, dgb, dgc):
        t, names = ctx.tensors, nat.LAYER_PARAM_FIELDS
        f_out, f_in = t["w_in"].shape
        num_gate = t["c_in"].numel()
        dev = t["w_in"].device
        k_ext = 3 * f_in + (f_in + 1 if ctx.has_res else 0) + 3
        dw_ext = torch.zeros((k_ext, f_out), device=dev) if dw_ext is None else dw_ext.contiguous()
        dga, dgb, dgc = (torch.zeros(num_gate, device=dev) if g is None else g.contiguous().reshape(-1) for g in (dga, dgb, dgc))
        grads = {k: (None if t[k] is None else torch.empty_like(t[k])) for k in names}
        p_ref, keep_p = nat.layer_params(**t)
        g_ref, keep_g = nat.layer_params(**grads)
        nat.call("pg_unpack_layer_param_grads", p_ref, nat.ptr(dw_ext), nat.ptr(dga), nat.ptr(dgb), nat.ptr(dgc), num_gate, f_in, f_out,
                 int(ctx.has_res), g_ref, nat.stream_ptr())
        out = [None if grads[k] is None else grads[k].reshape(shape) for k, shape in zip(names, ctx.shapes)]
        return (None, *out)


# ------------------------------------------------------------------------------------------------
# row f1: decoder output layer + log_softmax + nll_loss, forward and backward in one pass
# ------------------------------------------------------------------------------------------------
class _LinearLogSoftmaxNLL(torch.autograd.Function):
    """loss = nll_loss(log_softmax(d @ W^T + b), labels)   (reference protgram_directgcn.py:219-221 +
    protgram_directgcn_trainer.py:94).  The logits GEMM and the two gradient GEMMs are library GEMMs;
    everything between them (log_softmax, nll, their backward, the bias gradient: 10 passes over the
    N x C matrix in the reference) is one pass of pg_softmax_nll that overwrites the logits with
    d(loss)/d(logits).  labels outside [0, C) are ignored (ignore_index semantics)."""

    @staticmethod
    def forward(ctx, d, weight, bias, labels, has_ignored):
        nat.check_tensor(d, "decoder input")
        n, c = d.shape[0], weight.shape[0]
        dev = d.device
        k = d.shape[1]
        if TC_MODE != "off" and d.is_cuda and k % 4 == 0 and (TC_MODE == "force" or (n >= TC_MIN_ROWS and c >= TC_MIN_WIDTH)):
            # output layer on the tensor cores (3 x TF32): rows padded to a multiple of 4 columns, logits = a view of the buffer
            ld = (c + 3) // 4 * 4
            buf = torch.empty((n, ld), dtype=torch.float32, device=dev)
            ws_tc = nat.workspace(nat.query("pg_linear_tc_ws_bytes", k, c), dev)
            d32, w32 = d.contiguous().float(), weight.contiguous().float()
            nat.call("pg_linear_tc", nat.ptr(d32), d32.stride(0), n, k, nat.ptr(w32), nat.ptr(bias.contiguous().float()) if bias is not None else None,
                     c, nat.ptr(buf), ld, nat.ptr(ws_tc), ws_tc.numel(), nat.stream_ptr())
            logits = buf[:, :c]
        else:
            d32, w32 = d.contiguous().float(), weight.contiguous().float()
            logits = torch.empty((n, c), dtype=torch.float32, device=dev)
            nat.call("pg_linear_fwd", nat.ptr(d32), d32.stride(0), n, k, nat.ptr(w32), nat.ptr(bias.contiguous().float()) if bias is not None else None,
                     c, 0, nat.ptr(logits), logits.stride(0), nat.stream_ptr())
        # grad_scale is a host scalar of the C ABI: with every row counted (the trainer's case) it is 1/n and no
        # device->host read is needed; masked labels take the exact count from the device
        scale = 1.0 / max(n, 1)
        if has_ignored:
            scale = 1.0 / max(1, int(((labels >= 0) & (labels < c)).sum().item()))
        row_loss = torch.empty(n, dtype=torch.float32, device=dev)
        colsum = torch.empty(c, dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        ws = nat.workspace(nat.query("pg_softmax_nll_ws_bytes", n, c), dev)
        nat.call("pg_softmax_nll", nat.ptr(logits), logits.stride(0), n, c, nat.ptr(labels), scale, nat.ptr(row_loss),
                 nat.ptr(colsum), nat.ptr(loss), nat.ptr(ws), ws.numel(), nat.stream_ptr())
        ctx.save_for_backward(d, weight, logits, colsum)   # logits now holds dloss/dlogits
        ctx.has_bias = bias is not None
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        d, weight, g, colsum = ctx.saved_tensors
        dd = dw = db = None
        n, k = d.shape
        c = weight.shape[0]
        dev = d.device
        st = nat.stream_ptr()
        need_dd, need_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if (need_dd or need_dw) and nat.query("pg_decoder_grads_supported", k) and (c <= DECODER_GRADS_FUSED_MAX_CLASSES or DECODER_GRADS == "fused") \
                and DECODER_GRADS != "library":
            # both gradient GEMMs in ONE pass over the N x C gradient matrix (csrc/decoder.cu: decoder_grads_kernel)
            d32, w32 = d.contiguous().float(), weight.contiguous().float()
            dd = torch.empty((n, k), dtype=torch.float32, device=dev)
            dw = torch.empty((c, k), dtype=torch.float32, device=dev)
            ws = nat.workspace(nat.query("pg_decoder_grads_ws_bytes", n, c, k), dev)
            nat.call("pg_decoder_grads", nat.ptr(g), g.stride(0), nat.ptr(d32), d32.stride(0), nat.ptr(w32), n, c, k, 1.0, nat.ptr(dd),
                     nat.ptr(dw), nat.ptr(ws), ws.numel(), st)
            dd, dw = (dd.mul_(grad_out) if need_dd else None), (dw.mul_(grad_out) if need_dw else None)
        elif DECODER_GRADS == "simt":
            if need_dd:
                dd = _linear_bwd_data(g, weight, c, k).mul_(grad_out)
            if need_dw:
                dw = _linear_bwd_weight(g, d, c, k).mul_(grad_out)
        else:
            # C = N classes (next_node task, 8401^2 at config C2) with K = 32: two tall-skinny reductions over the 282 MB gradient
            # matrix.  Measured on B200 (profiles/r02_launches_c2_step.txt): library SGEMM pair 0.40 ms, decoder_grads_kernel<1>
            # 0.89 ms (shared-memory bound at 0.37 LDS per FMA), generic SIMT mainloop ~0.8 ms (half of its 64-wide tile idle at
            # K = 32) -- so this one shape stays on the library until the fused kernel is rebuilt on tensor cores.
            if need_dd:
                dd = (g @ weight).mul_(grad_out)
            if need_dw:
                dw = (g.t() @ d).mul_(grad_out)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = colsum * grad_out
        return dd, dw, db, None, None


def _linear_bwd_data(g: torch.Tensor, weight: torch.Tensor, c: int, k: int) -> torch.Tensor:
    """dx [N, K] = g [N, C] @ W [C, K] (SIMT fp32 mainloop, csrc/gemm.cu)."""
    n = g.shape[0]
    w32 = weight.contiguous().float()
    dx = torch.empty((n, k), dtype=torch.float32, device=g.device)
    nat.call("pg_linear_bwd_data", nat.ptr(g), g.stride(0), n, c, nat.ptr(w32), k, nat.ptr(dx), dx.stride(0), nat.stream_ptr())
    return dx


def _linear_bwd_weight(g: torch.Tensor, x: torch.Tensor, c: int, k: int) -> torch.Tensor:
    """dW [C, K] = g^T [C, N] @ x [N, K] (rows split over CTAs, fixed-order reduction)."""
    n = g.shape[0]
    x32 = x.contiguous().float()
    dw = torch.empty((c, k), dtype=torch.float32, device=g.device)
    ws = nat.workspace(nat.query("pg_linear_bwd_weight_ws_bytes", n, c, k), g.device)
    nat.call("pg_linear_bwd_weight", nat.ptr(g), g.stride(0), nat.ptr(x32), x32.stride(0), n, c, k, nat.ptr(dw), nat.ptr(ws), ws.numel(),
             nat.stream_ptr())
    return dw


class _Linear(torch.autograd.Function):
    """y = act(x W^T + b) through libpgb200 (the decoder MLP, reference :173-180): SIMT fp32 mainloop with bias / ReLU fused into the
    epilogue, or the tcgen05 kernel for wide outputs; backward = two GEMMs + fixed-order column sums.  No library GEMM."""

    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        nat.check_tensor(x, "decoder input")
        x32, w32 = x.contiguous().float(), weight.contiguous().float()
        b32 = bias.contiguous().float() if bias is not None else None
        n, k = x32.shape
        c = w32.shape[0]
        if (not relu and TC_MODE != "off" and k % 4 == 0 and (TC_MODE == "force" or (n >= TC_MIN_ROWS and c >= TC_MIN_WIDTH))):
            ld = (c + 3) // 4 * 4
            buf = torch.empty((n, ld), dtype=torch.float32, device=x.device)
            ws_tc = nat.workspace(nat.query("pg_linear_tc_ws_bytes", k, c), x.device)
            nat.call("pg_linear_tc", nat.ptr(x32), x32.stride(0), n, k, nat.ptr(w32), nat.ptr(b32), c, nat.ptr(buf), ld, nat.ptr(ws_tc),
                     ws_tc.numel(), nat.stream_ptr())
            out = buf[:, :c]
        else:
            out = torch.empty((n, c), dtype=torch.float32, device=x.device)
            nat.call("pg_linear_fwd", nat.ptr(x32), x32.stride(0), n, k, nat.ptr(w32), nat.ptr(b32), c, int(relu), nat.ptr(out), out.stride(0),
                     nat.stream_ptr())
        ctx.save_for_backward(x32, w32, out if relu else None)
        ctx.relu, ctx.has_bias = bool(relu), bias is not None
        return out

    @staticmethod
    def backward(ctx, gout):
        x32, w32, out = ctx.saved_tensors
        n, k = x32.shape
        c = w32.shape[0]
        g = gout.float()
        if ctx.relu:
            g = g * (out > 0)
        g = g.contiguous()
        dx = _linear_bwd_data(g, w32, c, k) if ctx.needs_input_grad[0] else None
        dw = _linear_bwd_weight(g, x32, c, k) if ctx.needs_input_grad[1] else None
        db = None
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = torch.empty(c, dtype=torch.float32, device=g.device)
            wsc = nat.workspace(nat.query("pg_colsum_ws_bytes", n, c), g.device)
            nat.call("pg_colsum", nat.ptr(g), g.stride(0), n, c, nat.ptr(db), nat.ptr(wsc), wsc.numel(), nat.stream_ptr())
        return dx, dw, db, None


def linear_log_softmax_nll(d: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], labels: torch.Tensor,
                           has_ignored: bool = False) -> torch.Tensor:
    """Fused `F.nll_loss(F.log_softmax(F.linear(d, weight, bias), -1), labels)` (mean over counted rows).
    Pass has_ignored=True when some labels lie outside [0, C) (costs one device->host read of the count)."""
    if weight.shape[0] > nat.SOFTMAX_NLL_MAX_CLASSES:
        return F.nll_loss(F.log_softmax(F.linear(d, weight, bias), dim=-1), labels)   # torch CUDA ops, no row cache that large
    return _LinearLogSoftmaxNLL.apply(d.contiguous(), weight.contiguous(), bias, labels.contiguous(), bool(has_ignored))


# ------------------------------------------------------------------------------------------------
# modules
# ------------------------------------------------------------------------------------------------
class DirectGCNLayer(nn.Module):
    """Same parameters / init as reference :26-91 (state_dict compatible)."""

    def __init__(self, in_channels: int, out_channels: int, num_nodes: int, use_vector_coeffs: bool = True):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.num_nodes = num_nodes
        self.use_vector_coeffs = use_vector_coeffs
        self.lin_main_in = nn.Linear(in_channels, out_channels, bias=False)
        self.lin_main_out = nn.Linear(in_channels, out_channels, bias=False)
        self.lin_undirected = nn.Linear(in_channels, out_channels, bias=False)
        self.bias_main_in = nn.Parameter(torch.Tensor(out_channels))
        self.bias_main_out = nn.Parameter(torch.Tensor(out_channels))
        self.bias_undirected = nn.Parameter(torch.Tensor(out_channels))
        self.lin_shared = nn.Linear(in_channels, out_channels, bias=False)
        self.bias_directed_shared_in = nn.Parameter(torch.Tensor(out_channels))
        self.bias_directed_shared_out = nn.Parameter(torch.Tensor(out_channels))
        self.bias_undirected_shared = nn.Parameter(torch.Tensor(out_channels))
        if self.use_vector_coeffs and self.num_nodes > 0:
            for name in ("C_in_vec", "C_out_vec", "C_directed_vec", "C_undirected_vec", "C_all_vec"):
                setattr(self, name, nn.Parameter(torch.Tensor(num_nodes, 1)))
        else:
            self.use_vector_coeffs = False
            for name in ("C_in", "C_out", "C_directed", "C_undirected", "C_all"):
                setattr(self, name, nn.Parameter(torch.Tensor(1)))
        self.constant = nn.Parameter(torch.Tensor(num_nodes, out_channels)) if self.num_nodes > 0 else None
        self.reset_parameters()

    def reset_parameters(self):
        for lin in (self.lin_main_in, self.lin_main_out, self.lin_shared, self.lin_undirected):
            nn.init.xavier_uniform_(lin.weight)
        for b in (self.bias_main_in, self.bias_main_out, self.bias_directed_shared_in, self.bias_directed_shared_out,
                  self.bias_undirected, self.bias_undirected_shared):
            nn.init.zeros_(b)
        sfx = "_vec" if self.use_vector_coeffs else ""
        for name in ("C_in", "C_out", "C_directed", "C_undirected", "C_all"):
            nn.init.ones_(getattr(self, name + sfx))
        if self.constant is not None:
            nn.init.xavier_uniform_(self.constant)

    # -- parameter packing (tiny torch ops; autograd splits dW_ext back onto the 4 Linears / 6 biases)
    def _gates(self, original_indices):
        sfx = "_vec" if self.use_vector_coeffs else ""
        c_in, c_out, c_dir, c_und, c_all = (getattr(self, k + sfx) for k in ("C_in", "C_out", "C_directed", "C_undirected", "C_all"))
        if self.use_vector_coeffs and original_indices is not None:
            c_in, c_out, c_dir, c_und, c_all = (t[original_indices] for t in (c_in, c_out, c_dir, c_und, c_all))
        cd = c_all * c_dir
        return (cd * c_in).reshape(-1), (cd * c_out).reshape(-1), (c_all * c_und).reshape(-1)

    def _constant_rows(self, original_indices):
        # reference :116-128: the constant is only added on the vector-coefficient path
        if not self.use_vector_coeffs or self.constant is None:
            return None
        return self.constant if original_indices is None else self.constant[original_indices]

    def _w_ext(self, res_weight=None, res_bias=None):
        ws = self.lin_shared.weight
        blocks = [(self.lin_main_in.weight + ws).t(), (self.lin_main_out.weight + ws).t(), (self.lin_undirected.weight + ws).t()]
        if res_weight is not None:
            blocks.append(res_weight.t())
        blocks.append(torch.stack([self.bias_main_in + self.bias_directed_shared_in,
                                   self.bias_main_out + self.bias_directed_shared_out,
                                   self.bias_undirected + self.bias_undirected_shared]))
        if res_weight is not None:
            rb = res_bias if res_bias is not None else torch.zeros(self.out_channels, device=ws.device, dtype=ws.dtype)
            blocks.append(rb.unsqueeze(0))
        return torch.cat(blocks, dim=0)

    def _packed(self, res_weight=None, res_bias=None):
        sfx = "_vec" if self.use_vector_coeffs else ""
        c = [getattr(self, k + sfx) for k in ("C_in", "C_out", "C_directed", "C_undirected", "C_all")]
        return _PackLayerParams.apply(res_weight is not None, self.lin_main_in.weight, self.lin_main_out.weight,
                                      self.lin_undirected.weight, self.lin_shared.weight, self.bias_main_in, self.bias_main_out,
                                      self.bias_undirected, self.bias_directed_shared_in, self.bias_directed_shared_out,
                                      self.bias_undirected_shared, res_weight, res_bias, *c)

    def _run(self, x, edges, original_indices, res_weight, res_bias, add_identity, slope):
        ei_in, ew_in, ei_out, ew_out, ei_und, ew_und = edges
        nat.check_tensor(x, "x")
        struct = get_structure((ei_in, ei_out, ei_und), (ew_in, ew_out, ew_und), x.shape[0])
        if original_indices is None or not self.use_vector_coeffs:
            w_ext, ga, gb, gc = self._packed(res_weight, res_bias)      # one kernel (csrc/params.cu)
        else:                                                             # cluster mini-batch: gates gathered per sub-graph node
            ga, gb, gc = self._gates(original_indices)
            w_ext = self._w_ext(res_weight, res_bias)
        return _DirectGCNFused.apply(x, ga, gb, gc, w_ext, self._constant_rows(original_indices),
                                     struct, res_weight is not None, add_identity, slope)

    def forward(self, x: torch.Tensor,
                edge_index_in: torch.Tensor, edge_weight_in: Optional[torch.Tensor],
                edge_index_out: torch.Tensor, edge_weight_out: Optional[torch.Tensor],
                edge_index_undirected: torch.Tensor, edge_weight_undirected: Optional[torch.Tensor],
                original_indices: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Reference :93-135 (no residual, no activation: those belong to the stack)."""
        return self._run(x, (edge_index_in, edge_weight_in, edge_index_out, edge_weight_out, edge_index_undirected,
                             edge_weight_undirected), original_indices, None, None, False, 1.0)


class ProtGramDirectGCN(nn.Module):
    """Reference :143-222."""

    LEAKY_SLOPE = 0.01  # F.leaky_relu default (reference :215)

    def __init__(self, layer_dims: List[int], num_graph_nodes: Optional[int], task_num_output_classes: int, n_gram_len: int,
                 one_gram_dim: int, max_pe_len: int, dropout: float, use_vector_coeffs: bool, l2_eps: float = 1e-12):
        super().__init__()
        self.n_gram_len = n_gram_len
        self.one_gram_dim = one_gram_dim
        self.dropout = dropout
        self.l2_eps = l2_eps
        self.pe_layer = None
        if one_gram_dim > 0 and max_pe_len > 0:
            self.pe_layer = nn.Embedding(max_pe_len, one_gram_dim)
        self.convs = nn.ModuleList()
        self.res_projs = nn.ModuleList()
        if not layer_dims or len(layer_dims) < 2:
            raise ValueError("layer_dims must contain at least input and output dimensions (length >= 2).")
        for i in range(len(layer_dims) - 1):
            in_dim, out_dim = layer_dims[i], layer_dims[i + 1]
            nodes = num_graph_nodes if num_graph_nodes is not None else 0
            self.convs.append(DirectGCNLayer(in_dim, out_dim, nodes, use_vector_coeffs and nodes > 0))
            self.res_projs.append(nn.Linear(in_dim, out_dim) if in_dim != out_dim else nn.Identity())
        final_dim = layer_dims[-1]
        hidden = final_dim // 2 if final_dim > 1 else 1
        self.decoder_fc = nn.Sequential(nn.Linear(final_dim, hidden), nn.ReLU(), nn.Dropout(p=0.5),
                                        nn.Linear(hidden, task_num_output_classes))

    def _apply_pe(self, x: torch.Tensor) -> torch.Tensor:
        if self.pe_layer is None:
            return x
        if self.n_gram_len > 0 and self.one_gram_dim > 0 and x.shape[1] == self.n_gram_len * self.one_gram_dim:
            k = min(self.n_gram_len, self.pe_layer.num_embeddings)
            if k > 0:
                xr = x.clone().view(-1, self.n_gram_len, self.one_gram_dim)
                xr[:, :k, :] += self.pe_layer.weight[:k].unsqueeze(0)
                return xr.view(-1, self.n_gram_len * self.one_gram_dim)
        return x

    def embed(self, data, return_layers: bool = False):
        """The layer stack only (PE + L fused layers) -> final hidden state h (reference :208-218).
        With return_layers=True also returns the list of per-layer activations."""
        x = getattr(data, "x", None)
        ei_in, ew_in = getattr(data, "edge_index_in", None), getattr(data, "edge_weight_in", None)
        ei_out, ew_out = getattr(data, "edge_index_out", None), getattr(data, "edge_weight_out", None)
        ei_und, ew_und = getattr(data, "edge_index_undirected_norm", None), getattr(data, "edge_weight_undirected_norm", None)
        original_indices = getattr(data, "original_indices", None)
        if x is None or ei_in is None or ei_out is None or ei_und is None:
            raise ValueError("ProtGramDirectGCN requires 'x', 'edge_index_in', 'edge_index_out', and "
                             "'edge_index_undirected_norm' in the Data object.")
        edges = (ei_in, ew_in, ei_out, ew_out, ei_und, ew_und)
        h = self._apply_pe(x)
        layers = []
        for conv, res in zip(self.convs, self.res_projs):
            if isinstance(res, nn.Linear):
                h = conv._run(h, edges, original_indices, res.weight, res.bias, False, self.LEAKY_SLOPE)
            else:
                h = conv._run(h, edges, original_indices, None, None, True, self.LEAKY_SLOPE)
            h = F.dropout(h, p=self.dropout, training=self.training)
            layers.append(h)
        return (h, layers) if return_layers else h

    def _decoder_hidden(self, h: torch.Tensor) -> torch.Tensor:
        """decoder_fc[0..2] of the reference (:173-176): Linear + ReLU (one kernel, fused epilogue) + Dropout(0.5)."""
        first, drop = self.decoder_fc[0], self.decoder_fc[2]
        return drop(_Linear.apply(h, first.weight, first.bias, True))

    def nll_loss(self, data, labels: torch.Tensor, has_ignored: bool = False) -> torch.Tensor:
        """`F.nll_loss(self(data)[0], labels)` of the reference trainer (protgram_directgcn_trainer.py:92-94)
        without materialising log_softmax: the decoder's output Linear, log_softmax and nll_loss (and their
        backward) run through the fused loss (row f1)."""
        h = self.embed(data)
        d = self._decoder_hidden(h)
        last = self.decoder_fc[-1]
        return linear_log_softmax_nll(d, last.weight, last.bias, labels, has_ignored)

    def forward(self, data) -> Tuple[torch.Tensor, torch.Tensor]:
        h = self.embed(data)
        last = self.decoder_fc[-1]
        # The reference's call pattern F.nll_loss(model(data)[0], y) needs the N x C log-probabilities as a tensor, so they are
        # materialised here (the decoder GEMMs run through libpgb200, log_softmax is torch's elementwise kernel); `nll_loss()`
        # above is the fused form that never writes them.
        task_logits = _Linear.apply(self._decoder_hidden(h), last.weight, last.bias, False)
        emb = EmbeddingProcessor.l2_normalize_torch(h, eps=self.l2_eps)
        return F.log_softmax(task_logits, dim=-1), emb
