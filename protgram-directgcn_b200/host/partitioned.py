"""Row-partitioned DirectGCN propagation over several GPUs (one process per GPU, torch.distributed).

SURVEY.md 8(e): the propagation path shards by rows with ONE exchange step per SpMM.  Rank r owns
the node rows [lo_r, hi_r) of the shared pattern (global int32 columns) and the matching slice of
X; W and the other parameters are replicated (the dense transform and its epilogue are row-local).

    forward   X_full = all_gather(X_local)            Z_local = fan-out SpMM(rows of r, X_full)
    backward  symmetric matrices (reference-built graphs):
                  G_full = all_gather(dZ_local)       dX_local = fan-in SpMM(rows of r, G_full)
              general matrices:
                  dX_part[N, F] = fan-in over the transposed local block, then reduce-scatter

The exchange is an NCCL all-gather of fp32 rows over NVLink (N*F*4*(g-1)/g bytes received per GPU);
power-law graphs touch almost every remote row, so gathering whole blocks beats per-row peer loads
(B200 peer LDG latency is ~3.5x local DRAM).  Rows are padded to ceil(N/g) so the collective is a
single all_gather_into_tensor.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from .. import _native as nat
from .protgram_directgcn import _Csr


def row_range(n: int, rank: int, world: int) -> Tuple[int, int, int]:
    """-> (lo, hi, rows_per_rank) with equal-size (padded) blocks."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per), per


def slice_rows(rowptr: torch.Tensor, col: torch.Tensor, vals: List[torch.Tensor], lo: int, hi: int, per: int) -> _Csr:
    """Local CSR of rows [lo, hi) (padded with empty rows up to `per`), columns stay global."""
    b, e = int(rowptr[lo]), int(rowptr[hi])
    rp = torch.full((per + 1,), e - b, dtype=torch.int64, device=rowptr.device)
    rp[: hi - lo + 1] = rowptr[lo:hi + 1] - b
    return _Csr(rp.contiguous(), col[b:e].contiguous(), [v[b:e].contiguous() for v in vals])


def _all_gather_rows(x_local: torch.Tensor, group) -> torch.Tensor:
    world = dist.get_world_size(group)
    out = torch.empty((world * x_local.shape[0], x_local.shape[1]), dtype=x_local.dtype, device=x_local.device)
    dist.all_gather_into_tensor(out, x_local.contiguous(), group=group)
    return out


def _reduce_scatter_rows(full: torch.Tensor, per: int, group) -> torch.Tensor:
    rank = dist.get_rank(group)
    if dist.get_backend(group) == "nccl":
        out = torch.empty((per, full.shape[1]), dtype=full.dtype, device=full.device)
        dist.reduce_scatter_tensor(out, full.contiguous(), group=group)
        return out
    dist.all_reduce(full, group=group)  # gloo (CPU tests) has no reduce_scatter
    return full[rank * per:(rank + 1) * per].clone()


def _spmm_fanout(csr: _Csr, x_full: torch.Tensor, rows: int, f: int) -> torch.Tensor:
    nv = len(csr.vals)
    z = torch.empty((rows, nv * f), dtype=torch.float32, device=x_full.device)
    v = csr.vals + [None] * (3 - nv)
    nat.call("pg_spmm_fanout", nat.ptr(csr.rowptr), nat.ptr(csr.col), nat.ptr(v[0]), nat.ptr(v[1]), nat.ptr(v[2]), nv, rows, f,
             nat.ptr(x_full), x_full.stride(0), nat.ptr(z), z.stride(0), 0, csr.plan(3 * f), nat.stream_ptr())
    return z


def _spmm_fanin(csr: _Csr, g_full: torch.Tensor, rows: int, f: int) -> torch.Tensor:
    nv = len(csr.vals)
    y = torch.empty((rows, f), dtype=torch.float32, device=g_full.device)
    v = csr.vals + [None] * (3 - nv)
    nat.call("pg_spmm_fanin", nat.ptr(csr.rowptr), nat.ptr(csr.col), nat.ptr(v[0]), nat.ptr(v[1]), nat.ptr(v[2]), nv, rows, f,
             nat.ptr(g_full), g_full.stride(0), 0, None, 0, nat.ptr(y), y.stride(0), 0, csr.plan(3 * f), nat.stream_ptr())
    return y


class _PartitionedFanout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_local, local_csr, local_csr_t, per, n_total, symmetric, group):
        nat.check_tensor(x_local, "x_local")
        f = x_local.shape[1]
        x_full = _all_gather_rows(x_local.float(), group)
        ctx.meta = (local_csr, local_csr_t, per, n_total, symmetric, group, f)
        return _spmm_fanout(local_csr, x_full, per, f)

    @staticmethod
    def backward(ctx, dz_local):
        local_csr, local_csr_t, per, n_total, symmetric, group, f = ctx.meta
        dz_local = dz_local.contiguous().float()
        if symmetric:
            g_full = _all_gather_rows(dz_local, group)
            dx = _spmm_fanin(local_csr, g_full, per, f)
        else:
            world = dist.get_world_size(group)
            part = _spmm_fanin(local_csr_t, dz_local, world * per, f)  # rows = global sources, cols = local targets
            dx = _reduce_scatter_rows(part, per, group)
        return dx, None, None, None, None, None, None


class RowPartitionedPropagation:
    """Z_local = [A_in X | A_out X | U X] for the rows this rank owns.

    Build once per graph from the FULL shared-pattern CSR (every rank holds or loads it, or builds
    its slice itself); `symmetric=True` is the reference-built case.  For general matrices pass
    `transposed` = the CSR grouped by source of the same local block (rows = global ids)."""

    def __init__(self, rowptr, col, vals, n: int, group=None, symmetric: bool = True, transposed: Optional[_Csr] = None):
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.n = n
        self.lo, self.hi, self.per = row_range(n, self.rank, self.world)
        self.local = slice_rows(rowptr, col, list(vals), self.lo, self.hi, self.per)
        self.symmetric = symmetric
        self.transposed = transposed
        if not symmetric and transposed is None:
            raise ValueError("general (non-symmetric) matrices need the source-grouped CSR of the local block")

    def pad_rows(self, x_local: torch.Tensor) -> torch.Tensor:
        if x_local.shape[0] == self.per:
            return x_local
        pad = torch.zeros((self.per - x_local.shape[0], x_local.shape[1]), dtype=x_local.dtype, device=x_local.device)
        return torch.cat([x_local, pad], dim=0)

    def __call__(self, x_local: torch.Tensor) -> torch.Tensor:
        """x_local: [hi-lo (or per), F] -> z_local [per, nv*F] (rows beyond hi-lo are zero)."""
        return _PartitionedFanout.apply(self.pad_rows(x_local), self.local, self.transposed, self.per, self.n,
                                        self.symmetric, self.group)
