"""Row-partitioned graphs over several GPUs (one process per GPU, torch.distributed).  Three parts:
  1. RowPartitionedPropagation  the SpMM with one exchange step (below)
  2. normalize_row_partitioned  the propagation matrices of the rank's rows from the rank's out-edges
  3. PartitionedStructure / partitioned_data / allreduce_replicated_grads  the unchanged model on a row block


SURVEY.md 8(e): the propagation path shards by rows with ONE exchange step per SpMM.  Rank r owns
the node rows [lo_r, hi_r) of the shared pattern and the matching slice of X; W and the other
parameters are replicated (the dense transform and its epilogue are row-local).

The exchange is a HALO exchange (north star: "all-gather of boundary features"): a rank's block only
references a subset of the remote rows -- on the randomly relabelled power-law graphs of config C5
about 30 % of them at 8 ranks (low-degree nodes are rarely anybody's neighbour), measured by
`HaloExchange.stats()` -- so each rank asks every peer ONCE, at structure-build time, for the sorted list
of rows it needs; per SpMM every rank packs the rows its peers asked for (`pg_gather_rows`) and one
`all_to_all_single` with those split sizes moves them over NVLink.  The block's columns are renumbered
once into [own rows | halo rows grouped by owner], the kernels read the two buffers through a split
operand (`pg_spmm_operand`), and the entries of a row keep their CSR order, so a block's result is
bitwise equal to the same rows of the single-GPU SpMM.

Overlap.  In the layer's BACKWARD the exchange of dY is posted first and the weight-gradient and gate-gradient GEMMs
(which need dY but no halo rows) run under it (`fanout_begin` / `fanout_finish`).  The forward has no independent work
to put there; the kernels can process the features in COLUMN CHUNKS so that chunk k+1's rows are in flight while the
SpMM works on chunk k (every output element is still summed in CSR order, `PGB200_EXCHANGE_CHUNKS`), but measured on
2 x B200 (R-MAT 2^22 nodes, F = 128) two 64-wide passes cost 2.5 ms more than one 128-wide pass and only 1.2 ms of
exchange is there to hide, so the default is ONE chunk; the chunked path stays for fabrics where the exchange dominates.

    forward   H_x = halo(X_local)                   Z_local = fan-out SpMM(rows of r, [X_local | H_x])
    backward  layer (symmetric matrices): dX = sum_v (A_v diag(g_v) dY) W_v^T -- the SAME exchange on the
              F_out-wide dY plus the three gate values per halo row (`PartitionedStructure.fanout(scales=...)`),
              never the 3F-wide gated gradient;
              bare operator Z = fan-out(X) given dZ:  halo(dZ_local) [3F wide] + local fan-in  (symmetric)
                                                      fan-in over the transposed block + reduce-scatter (general)
PGB200_EXCHANGE=allgather restores the round-1 exchange (all_gather_into_tensor of every row) for A/B timing.
"""
from __future__ import annotations

import os
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


EXCHANGE_MODE = os.environ.get("PGB200_EXCHANGE", "halo")          # "halo" | "allgather"
PIPELINE_CHUNKS = int(os.environ.get("PGB200_EXCHANGE_CHUNKS", "1"))  # feature-column chunks one exchange + SpMM is pipelined over (see module docstring: 1 = off)
PIPELINE_MIN_BYTES = 8 << 20                                         # below this a single exchange is cheaper than several


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


def halo_plan(col: torch.Tensor, lo: int, per: int, world: int):
    """Pure part of the halo exchange: from the (global, int32) columns of the block whose own rows are [lo, lo + per)
    -> (need: sorted global ids of the referenced remote rows = grouped by owner, need_counts[world], col_ext: the same
    entries renumbered into [own rows 0..per) | halo rows per.. in `need` order])."""
    c = col.to(torch.int64)
    own = (c >= lo) & (c < lo + per)
    need = torch.unique(c[~own])
    need_counts = torch.bincount(torch.div(need, max(per, 1), rounding_mode="floor"), minlength=world)[:world]
    ext = torch.where(own, c - lo, per + torch.searchsorted(need, c))
    return need, need_counts, ext.to(torch.int32).contiguous()


class HaloExchange:
    """Which remote rows this rank's block references, which of its own rows the peers reference, and the block's columns
    renumbered into [own rows (per) | halo rows (sorted by global id, i.e. grouped by owner)].  Built once per structure:
    two small all-to-alls (counts, row ids).  `exchange(t)` then moves the referenced rows of a [per, w] matrix."""

    def __init__(self, col: torch.Tensor, n: int, group=None):
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.n = int(n)
        self.lo, self.hi_row, self.per = row_range(n, self.rank, self.world)
        dev = col.device
        lo, per = self.lo, self.per
        need, need_counts, self.col_ext = halo_plan(col, lo, per, self.world)
        self.num_halo = int(need.numel())
        serve_counts = torch.empty_like(need_counts)
        dist.all_to_all_single(serve_counts, need_counts, group=self.group)
        self.need_splits = [int(v) for v in need_counts.tolist()]
        self.serve_splits = [int(v) for v in serve_counts.tolist()]
        serve = torch.empty(sum(self.serve_splits), dtype=torch.int64, device=dev)
        dist.all_to_all_single(serve, need, output_split_sizes=self.serve_splits, input_split_sizes=self.need_splits, group=self.group)
        self.serve_idx = (serve - lo).contiguous()                       # local row numbers, grouped by requesting peer
        self.num_serve = int(serve.numel())
        if self.num_serve and (int(self.serve_idx.min()) < 0 or int(self.serve_idx.max()) >= per):
            raise ValueError("halo exchange: a peer asked for a row outside this rank's block")
        self.need = need

    def close(self):
        """Release the peer-memory transport's buffers (collective; a no-op for the NCCL transport)."""
        tr = getattr(self, "_transport", None)
        if tr:
            tr.close()
        self._transport = False

    def stats(self) -> dict:
        remote = max(1, self.n - (self.hi_row - self.lo))
        return {"halo_rows": self.num_halo, "remote_rows": remote, "halo_fraction_of_remote_rows": self.num_halo / remote,
                "rows_served_to_peers": self.num_serve}

    def _peer_transport(self):
        if getattr(self, "_transport", None) is None:
            self._transport = False
            if (HALO_TRANSPORT == "p2p" and self.world > 1 and dist.get_backend(self.group) == "nccl" and self.serve_idx.is_cuda):
                self._transport = PeerMemoryTransport(self)
        return self._transport or None

    def start(self, t: torch.Tensor, async_op: bool = True):
        """t: [per, w] fp32 (any row stride).  Moves the rows the peers asked for: the peer-memory push kernel on an NVLink
        node (PeerMemoryTransport), else pack + all_to_all_single.  -> (halo buffer [num_halo, w], work handle or None, keep-alive)."""
        tr = self._peer_transport()
        if tr is not None and tr.fits(int(t.shape[1])) and tr._slots_for(int(t.shape[1])) is not None:
            recv, work, keep = tr.start(t)
            if not async_op:
                work.wait()
                work = None
            return recv, work, keep
        w = int(t.shape[1])
        send = torch.empty((self.num_serve, w), dtype=torch.float32, device=t.device)
        if self.num_serve:
            nat.call("pg_gather_rows", nat.ptr(t), t.stride(0), nat.ptr(self.serve_idx), self.num_serve, w, nat.ptr(send), w, nat.stream_ptr())
        recv = torch.empty((self.num_halo, w), dtype=torch.float32, device=t.device)
        work = dist.all_to_all_single(recv, send, output_split_sizes=self.need_splits, input_split_sizes=self.serve_splits,
                                      group=self.group, async_op=async_op)
        return recv, (work if async_op else None), send

    def exchange(self, t: torch.Tensor) -> torch.Tensor:
        recv, _, _ = self.start(t, async_op=False)
        return recv


HALO_TRANSPORT = os.environ.get("PGB200_HALO_TRANSPORT", "p2p")    # "p2p": peer-memory push kernel (NCCL groups on one node) | "nccl"


class _ReadyWork:
    """`work.wait()` of the peer-memory transport: enqueue the one-CTA wait kernel for this exchange's epoch on the current stream."""

    def __init__(self, transport, epoch):
        self.transport, self.epoch = transport, epoch

    def wait(self):
        t = self.transport
        nat.call("pg_halo_wait", t.ctrl_ptr, t.world, t.rank, self.epoch, nat.ptr(t.err), nat.stream_ptr())


class PeerMemoryTransport:
    """The halo exchange as ONE kernel over NVLink peer memory (csrc/peer.cu) instead of pack + all_to_all_single.
    Every rank owns, per exchanged width, a ring of RING receive slots that its peers map through CUDA IPC; `start` launches
    pg_halo_push on a side stream (rows read straight out of the operand, stored into the peers' current slot, epoch flag
    published when the stores are fenced) and returns the local slot as a tensor plus a handle whose `wait()` puts
    pg_halo_wait in front of the consumer.  Setting it up is collective (IPC handles are all-gathered)."""

    RING = 4
    MAX_RING_BYTES = 40 << 30      # wider exchanges (the bare operator's 3F-wide gradient at C5 size) stay on all_to_all_single

    def fits(self, w: int) -> bool:
        """Same answer on every rank (decided from the all-gathered halo sizes): is the receive ring of this width affordable?"""
        return self.RING * max(1, max(self.peer_halo_rows)) * int(w) * 4 <= self.MAX_RING_BYTES

    def check(self):
        """Raise if a wait kernel ever gave up (host sync)."""
        code = int(self.err.item())
        if code:
            raise nat.NativeError(f"halo exchange: peer {code - 1} never published its epoch (pg_halo_wait gave up)")

    def __init__(self, halo: "HaloExchange"):
        self.halo, self.group, self.rank, self.world = halo, halo.group, halo.rank, halo.world
        dev = halo.serve_idx.device
        self.dev = dev
        if self.world > nat.query_const("PG_MAX_PEERS"):
            raise ValueError("peer-memory halo transport supports up to PG_MAX_PEERS ranks")
        # who needs how many rows from whom: counts[p][q] = rows rank p receives from rank q
        mine = torch.tensor(halo.need_splits, dtype=torch.int64, device=dev)
        allc = torch.empty((self.world, self.world), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allc, mine, group=self.group)
        counts = allc.tolist()
        self.peer_halo_rows = [int(sum(counts[p])) for p in range(self.world)]
        self.my_offset_on_peer = [int(sum(counts[p][:self.rank])) for p in range(self.world)]   # first row of MY rows in peer p's halo order
        self.row_begin = [0]
        for n_rows in halo.serve_splits:
            self.row_begin.append(self.row_begin[-1] + int(n_rows))
        # control block: one flag word per sender
        self.ctrl_ptr, handle = nat.peer_alloc(256)
        self.peer_ctrl = self._exchange_and_open(handle, self.ctrl_ptr)
        self.done = torch.zeros(1, dtype=torch.int32, device=dev)
        self.err = torch.zeros(1, dtype=torch.int32, device=dev)
        self.epoch = 0
        self.slots = {}
        self.stream = torch.cuda.Stream(device=dev)
        self._keep = []

    def _exchange_and_open(self, handle: bytes, own_ptr: int):
        h = torch.frombuffer(bytearray(handle), dtype=torch.uint8).to(self.dev)
        allh = torch.empty((self.world, 64), dtype=torch.uint8, device=self.dev)
        dist.all_gather_into_tensor(allh, h, group=self.group)
        allh = allh.cpu().numpy()
        return [own_ptr if p == self.rank else nat.peer_open(bytes(allh[p].tobytes())) for p in range(self.world)]

    def _slots_for(self, w: int):
        """Receive ring of one width; None when some rank could not allocate it (then EVERY rank uses all_to_all_single for this
        width).  Collective: every rank meets the widths in the same order (same model, same calls)."""
        if w not in self.slots:
            rows = max(1, self.halo.num_halo)
            torch.cuda.empty_cache()                       # cudaMalloc competes with torch's cached blocks
            local_ptr, handle = 0, bytes(64)
            try:
                local_ptr, handle = nat.peer_alloc(self.RING * rows * w * 4)
            except nat.NativeError:
                local_ptr = 0
            ok = torch.tensor([1 if local_ptr else 0], dtype=torch.int32, device=self.dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
            if int(ok.item()) == 0:
                if local_ptr:
                    nat.call("pg_peer_free", local_ptr)
                self.slots[w] = None
            else:
                self.slots[w] = (local_ptr, self._exchange_and_open(handle, local_ptr))
        return self.slots[w]

    def close(self):
        """Collective teardown: unmap the peers' buffers, then free the own ones (no rank may still be exchanging)."""
        if self.ctrl_ptr == 0:
            return
        torch.cuda.synchronize(self.dev)
        dist.barrier(group=self.group)
        own = [self.ctrl_ptr] + [v[0] for v in self.slots.values() if v]
        for ptrs in [self.peer_ctrl] + [v[1] for v in self.slots.values() if v]:
            for p, ptr in enumerate(ptrs):
                if p != self.rank:
                    nat.call("pg_peer_close", ptr)
        dist.barrier(group=self.group)
        for ptr in own:
            nat.call("pg_peer_free", ptr)
        self.ctrl_ptr, self.slots = 0, {}

    def start(self, t: torch.Tensor):
        """t: [per, w] fp32.  -> (receive slot [num_halo, w] as a tensor, wait handle, keep-alive)."""
        import ctypes
        w = int(t.shape[1])
        local_ptr, peer_ptrs = self._slots_for(w)
        self.epoch += 1
        if self.epoch % 256 == 0 and int(self.err.item()):
            raise nat.NativeError(f"halo exchange: peer {int(self.err.item()) - 1} never published its epoch (pg_halo_wait gave up)")
        slot = self.epoch % self.RING
        P = ctypes.c_void_p
        dst = (P * self.world)(*[P(peer_ptrs[p] + (slot * max(1, self.peer_halo_rows[p]) + self.my_offset_on_peer[p]) * w * 4)
                                 if p != self.rank and self.halo.serve_splits[p] else P(None) for p in range(self.world)])
        flag = (P * self.world)(*[P(self.peer_ctrl[p] + 4 * self.rank) if p != self.rank else P(None) for p in range(self.world)])
        rb = (ctypes.c_int64 * (self.world + 1))(*self.row_begin)
        cur = torch.cuda.current_stream(self.dev)
        self.stream.wait_stream(cur)                       # the operand is complete
        with torch.cuda.stream(self.stream):
            nat.call("pg_halo_push", nat.ptr(t), t.stride(0), nat.ptr(self.halo.serve_idx), rb, dst, flag, self.world, self.rank, w, w, self.epoch,
                     nat.ptr(self.done), nat.stream_ptr())
        t.record_stream(self.stream)
        recv = nat.view_tensor(local_ptr + slot * max(1, self.halo.num_halo) * w * 4, (self.halo.num_halo, w), self.dev)
        return recv, _ReadyWork(self, self.epoch), t


def _feature_chunks(f: int, per: int, world: int) -> List[Tuple[int, int]]:
    """(first column, width) of the feature-column chunks one exchange + SpMM is pipelined over (widths are multiples of 4:
    the kernels' 128-bit path)."""
    k = PIPELINE_CHUNKS if (world > 1 and f % 4 == 0 and per * f * 4 >= PIPELINE_MIN_BYTES) else 1
    k = max(1, min(k, f // 4))
    if k == 1:
        return [(0, f)]
    w = -(-(f // 4) // k) * 4
    return [(c0, min(w, f - c0)) for c0 in range(0, f, w)]


class _PendingFanout:
    """A fan-out whose halo exchange has been posted (NCCL's stream) but whose SpMM has not been launched yet: whatever the
    caller enqueues between `_halo_fanout_begin` and `_halo_fanout_finish` runs UNDER the exchange (the layer's backward puts
    its weight-gradient and gate-gradient GEMMs there, which need dY but no halo rows)."""
    __slots__ = ("csr_ext", "halo", "x", "f", "chunks", "posted", "mine", "s_halo", "s_work", "s_keep", "scales", "scale_stride")


def _halo_fanout_begin(csr_ext: _Csr, halo: HaloExchange, x: torch.Tensor, f: int, scales=None, scale_stride: int = 1) -> _PendingFanout:
    p = _PendingFanout()
    p.csr_ext, p.halo, p.f, p.scales, p.scale_stride = csr_ext, halo, f, scales, scale_stride
    p.x = x.contiguous()
    per = halo.per
    p.mine = p.s_halo = p.s_work = p.s_keep = None
    if scales is not None and scale_stride == 1:     # per-node gates: the kernel scales by the SOURCE row -> the halo rows' gates travel too
        p.mine = torch.zeros((per, 4), dtype=torch.float32, device=x.device)
        for k, sc in enumerate(scales):
            p.mine[:, k] = sc.reshape(-1)
        p.s_halo, p.s_work, p.s_keep = halo.start(p.mine)
    p.chunks = _feature_chunks(f, per, halo.world)
    tr = halo._peer_transport()
    if tr is not None and len(p.chunks) + (1 if p.mine is not None else 0) > tr.RING - 2:
        p.chunks = [(0, f)]      # the receive ring holds two exchanges per consume phase safely (see PeerMemoryTransport); no deeper pipelining there
    p.posted = [halo.start(p.x[:, c0:c0 + w]) for c0, w in p.chunks]       # all packs + exchanges are queued up front
    return p


def _halo_fanout_finish(p: _PendingFanout) -> torch.Tensor:
    csr_ext, halo, x, f = p.csr_ext, p.halo, p.x, p.f
    per, nv = halo.per, len(csr_ext.vals)
    z = torch.empty((per, nv * f), dtype=torch.float32, device=x.device)
    v = csr_ext.vals + [None] * (3 - nv)
    s_ptrs, s_stride = (None, None, None), 0
    if p.scales is not None:
        if p.scale_stride == 1:
            if p.s_work is not None:
                p.s_work.wait()
            s_ext = torch.cat([p.mine, p.s_halo], dim=0).reshape(-1)    # [(per + H) * 4]: gates of own rows, then of the halo rows
            s_ptrs, s_stride = tuple(nat.ptr(s_ext[k:]) for k in range(3)), 4
        else:
            s_ptrs, s_stride = tuple(nat.ptr(sc) for sc in p.scales), 0
    for (c0, w), (recv, work, _keep) in zip(p.chunks, p.posted):
        if work is not None:
            work.wait()                                               # the compute stream waits for THIS chunk only
        nat.call("pg_spmm_fanout_split", nat.ptr(csr_ext.rowptr), nat.ptr(csr_ext.col), nat.ptr(v[0]), nat.ptr(v[1]), nat.ptr(v[2]), nv,
                 per, w, nat.spmm_operand(x[:, c0:c0 + w], recv, per), nat.ptr(z[:, c0:]), z.stride(0), 0, f,
                 s_ptrs[0], s_ptrs[1] if nv == 3 else None, s_ptrs[2] if nv == 3 else None, s_stride, csr_ext.plan(3 * f), nat.stream_ptr())
    return z


def _halo_fanout(csr_ext: _Csr, halo: HaloExchange, x: torch.Tensor, f: int, scales=None, scale_stride: int = 1) -> torch.Tensor:
    """Z[per, nv f] = fan-out SpMM of this rank's rows over [x | halo rows of x].  scales (backward of the gated layer):
    per-source-row gates; with scale_stride == 1 their halo values travel as one extra 4-float exchange."""
    return _halo_fanout_finish(_halo_fanout_begin(csr_ext, halo, x, f, scales, scale_stride))


def _halo_fanin(csr_ext: _Csr, halo: HaloExchange, dz: torch.Tensor, f: int, init: Optional[torch.Tensor]) -> torch.Tensor:
    """dX[per, f] = (init) + sum_v A_v[rows of this rank] dZ_v for symmetric matrices: halo rows of dZ (nv f wide) + local fan-in."""
    per, nv = halo.per, len(csr_ext.vals)
    dz = dz.contiguous()
    recv = halo.exchange(dz)
    dx = torch.empty((per, f), dtype=torch.float32, device=dz.device)
    v = csr_ext.vals + [None] * (3 - nv)
    nat.call("pg_spmm_fanin_split", nat.ptr(csr_ext.rowptr), nat.ptr(csr_ext.col), nat.ptr(v[0]), nat.ptr(v[1]), nat.ptr(v[2]), nv, per, f,
             nat.spmm_operand(dz, recv, per), 0, f, nat.ptr(init), init.stride(0) if init is not None else 0, nat.ptr(dx), dx.stride(0), 0,
             csr_ext.plan(3 * f), nat.stream_ptr())
    return dx


def _fanin_exchanged(csr: _Csr, dz_local: torch.Tensor, per: int, f: int, init: Optional[torch.Tensor], group) -> torch.Tensor:
    """Round-1 exchange (PGB200_EXCHANGE=allgather): all-gather of dZ [N, 3F] + local fan-in over global columns."""
    nv = len(csr.vals)
    g_full = _all_gather_rows(dz_local.contiguous(), group)
    dx = torch.empty((per, f), dtype=torch.float32, device=dz_local.device)
    v = csr.vals + [None] * (3 - nv)
    nat.call("pg_spmm_fanin", nat.ptr(csr.rowptr), nat.ptr(csr.col), nat.ptr(v[0]), nat.ptr(v[1]), nat.ptr(v[2]), nv, per, f,
             nat.ptr(g_full), g_full.stride(0), 0, nat.ptr(init), init.stride(0) if init is not None else 0, nat.ptr(dx), dx.stride(0), 0,
             csr.plan(3 * f), nat.stream_ptr())
    return dx


class _PartitionedFanout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_local, prop):
        nat.check_tensor(x_local, "x_local")
        f = x_local.shape[1]
        ctx.prop, ctx.f = prop, f
        x_local = x_local.float()
        if prop.halo is not None:
            return _halo_fanout(prop.local_ext, prop.halo, x_local, f)
        return _spmm_fanout(prop.local, _all_gather_rows(x_local, prop.group), prop.per, f)

    @staticmethod
    def backward(ctx, dz_local):
        prop, f = ctx.prop, ctx.f
        dz_local = dz_local.contiguous().float()
        if prop.symmetric and prop.halo is not None:
            dx = _halo_fanin(prop.local_ext, prop.halo, dz_local, f, None)
        elif prop.symmetric:
            dx = _fanin_exchanged(prop.local, dz_local, prop.per, f, None, prop.group)
        else:
            part = _spmm_fanin(prop.transposed, dz_local, prop.world * prop.per, f)  # rows = global sources, cols = local targets
            dx = _reduce_scatter_rows(part, prop.per, prop.group)
        return dx, None


class RowPartitionedPropagation:
    """Z_local = [A_in X | A_out X | U X] for the rows this rank owns.

    Build once per graph from the FULL shared-pattern CSR (every rank holds or loads it, or builds
    its slice itself); `symmetric=True` is the reference-built case.  For general matrices pass
    `transposed` = the CSR grouped by source of the same local block (rows = global ids)."""

    def __init__(self, rowptr, col, vals, n: int, group=None, symmetric: bool = True, transposed: Optional[_Csr] = None):
        group = group if group is not None else dist.group.WORLD
        lo, hi, per = row_range(n, dist.get_rank(group), dist.get_world_size(group))
        self._init(slice_rows(rowptr, col, list(vals), lo, hi, per), n, group, symmetric, transposed)

    def _init(self, local: _Csr, n: int, group, symmetric: bool, transposed: Optional[_Csr]):
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.n = n
        self.lo, self.hi, self.per = row_range(n, self.rank, self.world)
        if local.rowptr.numel() != self.per + 1:
            raise ValueError(f"local rowptr must hold {self.per + 1} entries (padded block), got {local.rowptr.numel()}")
        self.local, self.symmetric, self.transposed = local, symmetric, transposed
        if not symmetric and transposed is None:
            raise ValueError("general (non-symmetric) matrices need the source-grouped CSR of the local block")
        self.halo = self.local_ext = None
        if EXCHANGE_MODE == "halo":
            self.halo = HaloExchange(local.col, n, self.group)
            self.local_ext = _Csr(local.rowptr, self.halo.col_ext, local.vals)
            self.local_ext._plan = local._plan

    def pad_rows(self, x_local: torch.Tensor) -> torch.Tensor:
        if x_local.shape[0] == self.per:
            return x_local
        pad = torch.zeros((self.per - x_local.shape[0], x_local.shape[1]), dtype=x_local.dtype, device=x_local.device)
        return torch.cat([x_local, pad], dim=0)

    def __call__(self, x_local: torch.Tensor) -> torch.Tensor:
        """x_local: [hi-lo (or per), F] -> z_local [per, nv*F] (rows beyond hi-lo are zero)."""
        return _PartitionedFanout.apply(self.pad_rows(x_local), self)

    @classmethod
    def from_local(cls, local: _Csr, n: int, group=None, symmetric: bool = True, transposed: Optional[_Csr] = None):
        """From a row block that already lives on this rank (`normalize_row_partitioned`): rowptr has
        `per` + 1 entries (rows past the block are empty), columns are global."""
        self = cls.__new__(cls)
        self._init(local, n, group, symmetric, transposed)
        return self


# ----------------------------------------------------------------------------------------------
# Row-partitioned normalisation (SURVEY.md 8(e), row "Normalisation (a7-a9)"; reference
# graph_utils.py:140-287).  Rank r holds the coalesced out-edges of its rows [lo_r, hi_r).  The
# propagation matrices of a row need the transposed entries too, so there is exactly ONE exchange of
# edges (every edge i -> j goes to the owner of j) plus all-gathers of three per-node vectors:
#
#     all_to_all   (src, dst, w) by owner(dst)                  16 + 4 B per edge that leaves the rank
#     all_gather   weighted out-/in-degree (fp64)               16 B per node
#     all_gather   undirected degree (int32)                     4 B per node
#
# The local work is the row-block variant of the single-GPU kernels (csrc/graph_rows.cu); each
# block is bitwise equal to the same rows of DirectedNgramGraph's matrices built on one GPU.
# ----------------------------------------------------------------------------------------------
class RowBlockNormalizer:
    """Local phases for one row block; the caller moves the per-node vectors between them."""

    def __init__(self, o_src, o_dst, o_w, i_src, i_dst, i_w, n: int, lo: int, hi: int, eps: float = 1e-9):
        for t in (o_src, o_dst, o_w, i_src, i_dst, i_w):
            nat.check_tensor(t, "edge array")
        self.o_src, self.o_dst, self.o_w = o_src.contiguous(), o_dst.contiguous(), o_w.contiguous().float()
        self.i_src, self.i_dst, self.i_w = i_src.contiguous(), i_dst.contiguous(), i_w.contiguous().float()
        self.n, self.lo, self.rows, self.eps = int(n), int(lo), int(hi - lo), float(eps)
        self.nnz_o, self.nnz_i = int(o_src.numel()), int(i_src.numel())
        self.dev = o_src.device
        self.ws = None

    def degree_sums(self) -> Tuple[torch.Tensor, torch.Tensor]:
        rs_out = torch.zeros(self.rows, dtype=torch.float64, device=self.dev)
        rs_in = torch.zeros(self.rows, dtype=torch.float64, device=self.dev)
        if self.rows:
            nat.call("pg_degree_sums_rows", nat.ptr(self.o_src), nat.ptr(self.o_w), self.nnz_o, nat.ptr(self.i_dst), nat.ptr(self.i_w),
                     self.nnz_i, self.lo, self.rows, nat.ptr(rs_out), nat.ptr(rs_in), nat.stream_ptr())
        return rs_out, rs_in

    def structure(self) -> torch.Tensor:
        """Sort + pattern of the block -> its undirected degrees (int32[rows]) for the all-gather."""
        dev, rows = self.dev, self.rows
        if rows == 0:
            self.p = 0
            self.rowptr = torch.zeros(1, dtype=torch.int64, device=dev)
            self.col = torch.empty(0, dtype=torch.int32, device=dev)
            self.ain = tuple(torch.empty(0, dtype=dt, device=dev) for dt in (torch.int64, torch.int64, torch.float32))
            return torch.empty(0, dtype=torch.int32, device=dev)
        self.ws = nat.workspace(nat.query("pg_normalize_rows_ws_bytes", self.nnz_o, self.nnz_i, rows), dev)
        sizes = torch.zeros(2, dtype=torch.int64, device=dev)
        st = nat.stream_ptr()
        nat.call("pg_normalize_rows_sizes", nat.ptr(self.o_src), nat.ptr(self.o_dst), self.nnz_o, nat.ptr(self.i_src), nat.ptr(self.i_dst),
                 self.nnz_i, self.n, self.lo, rows, nat.ptr(sizes), nat.ptr(self.ws), self.ws.numel(), st)
        p, bad = (int(v) for v in sizes.tolist())
        if bad:
            raise ValueError(f"row block [{self.lo}, {self.lo + rows}): an out-edge's source or an in-edge's target lies outside "
                             "the block, or a node id is >= num_nodes")
        self.p = p
        self.rowptr = torch.empty(rows + 1, dtype=torch.int64, device=dev)
        self.col = torch.empty(p, dtype=torch.int32, device=dev)
        self.native = torch.zeros(rows, dtype=torch.uint8, device=dev)
        self.ain = (torch.empty(self.nnz_i, dtype=torch.int64, device=dev), torch.empty(self.nnz_i, dtype=torch.int64, device=dev),
                    torch.empty(self.nnz_i, dtype=torch.float32, device=dev))
        nat.call("pg_normalize_rows_structure", nat.ptr(self.i_w), self.nnz_o, self.nnz_i, self.n, self.lo, rows, p, nat.ptr(self.ain[0]),
                 nat.ptr(self.ain[1]), nat.ptr(self.ain[2]), nat.ptr(self.rowptr), nat.ptr(self.col), nat.ptr(self.native),
                 nat.ptr(self.ws), self.ws.numel(), st)
        return ((self.rowptr[1:] - self.rowptr[:-1]) + self.native).to(torch.int32)

    def values(self, rs_out: torch.Tensor, rs_in: torch.Tensor, deg: torch.Tensor) -> dict:
        """rs_out / rs_in (fp64) and deg (int32) are the GLOBAL [n] vectors."""
        dev, p = self.dev, self.p
        v_out, v_in, v_und = (torch.empty(p, dtype=torch.float32, device=dev) for _ in range(3))
        if self.rows:
            nat.call("pg_normalize_rows_values", nat.ptr(self.o_w), nat.ptr(self.i_w), self.nnz_o, self.nnz_i, self.n, self.lo, self.rows,
                     nat.ptr(rs_out), nat.ptr(rs_in), nat.ptr(deg), nat.ptr(self.native), self.eps, nat.ptr(v_out), nat.ptr(v_in),
                     nat.ptr(v_und), nat.ptr(self.ws), self.ws.numel(), nat.stream_ptr())
        self.ws = None
        return {"rowptr": self.rowptr, "col": self.col, "val_out": v_out, "val_in": v_in, "val_und": v_und, "pattern_nnz": p,
                "in_src": self.ain[0], "in_dst": self.ain[1], "in_w": self.ain[2]}


def exchange_in_edges(src: torch.Tensor, dst: torch.Tensor, w: torch.Tensor, n: int, group=None, by: str = "dst"):
    """Every rank passes the out-edges of its rows; returns the edges whose TARGET it owns
    (src, dst, w; one all-to-all of 20 B per edge).  Stable: edges from one peer keep their order.
    by="src" sends every edge to the owner of its SOURCE instead (re-dealing an edge table that is
    partitioned some other way, e.g. by key range, onto the equal row blocks of `row_range`)."""
    group = group if group is not None else dist.group.WORLD
    world = dist.get_world_size(group)
    per = max(1, (n + world - 1) // world)
    e = int(src.numel())
    dev = src.device
    owner = torch.div(dst if by == "dst" else src, per, rounding_mode="floor").contiguous()
    perm = torch.arange(e, dtype=torch.int32, device=dev)
    if e:
        # stable partition by owner: one radix pass of the library's own sort
        alt_k, alt_v = torch.empty_like(owner), torch.empty_like(perm)
        ws = nat.workspace(nat.query("pg_sort_pairs_ws_bytes", e), dev)
        nat.call("pg_sort_pairs", nat.ptr(owner), nat.ptr(alt_k), nat.ptr(perm), nat.ptr(alt_v), e, max(1, (world - 1).bit_length()),
                 nat.ptr(ws), ws.numel(), nat.stream_ptr())
    send_counts = torch.bincount(owner, minlength=world)[:world]
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    s_split, r_split = send_counts.tolist(), recv_counts.tolist()
    perm = perm.long()
    send_sd = torch.stack([src[perm], dst[perm]], dim=1).contiguous()
    send_w = w[perm].contiguous()
    total = int(sum(r_split))
    recv_sd = torch.empty((total, 2), dtype=torch.int64, device=dev)
    recv_w = torch.empty(total, dtype=w.dtype, device=dev)
    dist.all_to_all_single(recv_sd, send_sd, output_split_sizes=r_split, input_split_sizes=s_split, group=group)
    dist.all_to_all_single(recv_w, send_w, output_split_sizes=r_split, input_split_sizes=s_split, group=group)
    return recv_sd[:, 0].contiguous(), recv_sd[:, 1].contiguous(), recv_w


def _all_gather_vector(local: torch.Tensor, per: int, n: int, group) -> torch.Tensor:
    world = dist.get_world_size(group)
    padded = torch.zeros(per, dtype=local.dtype, device=local.device)
    padded[: local.numel()] = local
    out = torch.empty(world * per, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return out[:n].contiguous()


def normalize_row_partitioned(src: torch.Tensor, dst: torch.Tensor, w: torch.Tensor, n: int, eps: float = 1e-9, group=None) -> dict:
    """Propagation matrices of a row-partitioned graph.  `src, dst, w`: the coalesced out-edges of
    this rank's rows (`row_range(n, rank, world)`), int64 / int64 / fp32 on the GPU.  Returns the
    block in the layer's format (rowptr int64[per+1] padded with empty rows, col int32 global,
    val_out / val_in / val_und) plus the block's rows of A_in_w; `local_csr()` of the result feeds
    `RowPartitionedPropagation.from_local`."""
    group = group if group is not None else dist.group.WORLD
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi, per = row_range(n, rank, world)
    total = torch.tensor([int(src.numel())], dtype=torch.int64, device=src.device)
    dist.all_reduce(total, group=group)
    if int(total.item()) == 0:
        # no edge anywhere: the reference leaves all matrices EMPTY (graph_utils.py:219-223 and the constructor's else branch), not the identity
        dev = src.device
        e_i, e_f = torch.empty(0, dtype=torch.int64, device=dev), torch.empty(0, dtype=torch.float32, device=dev)
        return {"rowptr": torch.zeros(per + 1, dtype=torch.int64, device=dev), "col": torch.empty(0, dtype=torch.int32, device=dev),
                "val_out": e_f, "val_in": e_f.clone(), "val_und": e_f.clone(), "pattern_nnz": 0, "in_src": e_i, "in_dst": e_i.clone(),
                "in_w": e_f.clone(), "lo": lo, "hi": hi, "per": per, "rs_out": torch.zeros(n, dtype=torch.float64, device=dev),
                "rs_in": torch.zeros(n, dtype=torch.float64, device=dev), "deg": torch.zeros(n, dtype=torch.int32, device=dev)}
    i_src, i_dst, i_w = exchange_in_edges(src, dst, w, n, group)
    blk = RowBlockNormalizer(src, dst, w, i_src, i_dst, i_w, n, lo, hi, eps)
    rs_out_l, rs_in_l = blk.degree_sums()
    rs_out = _all_gather_vector(rs_out_l, per, n, group)
    rs_in = _all_gather_vector(rs_in_l, per, n, group)
    deg = _all_gather_vector(blk.structure(), per, n, group)
    res = blk.values(rs_out, rs_in, deg)
    rp = torch.full((per + 1,), res["pattern_nnz"], dtype=torch.int64, device=src.device)
    rp[: hi - lo + 1] = res["rowptr"]
    res.update(rowptr=rp, lo=lo, hi=hi, per=per, rs_out=rs_out, rs_in=rs_in, deg=deg)
    return res


def local_csr(res: dict) -> _Csr:
    """[val_in, val_out, val_und] in the order the layer consumes them (Z = [A_in X | A_out X | U X])."""
    return _Csr(res["rowptr"], res["col"], [res["val_in"], res["val_out"], res["val_und"]])


# ----------------------------------------------------------------------------------------------
# Row-partitioned DirectGCN MODEL (SURVEY.md 8(e), row "Propagation (a10-a12)"): the unchanged
# ProtGramDirectGCN / DirectGCNLayer classes run on the rows of one rank.  Everything in a layer is
# row-local (gates, dense transform, bias, constant, residual, activation, decoder, log_softmax, L2
# normalisation) except the three SpMMs, which see all rows: the structure object below answers the
# layer's fan-out / fan-in calls with an all-gather of the operand + the local kernels.  Weights and
# biases are replicated (their gradients are partial sums: `allreduce_replicated_grads`); the
# per-node parameters (`constant`, the gate vectors) exist for the local rows only.
# ----------------------------------------------------------------------------------------------
class PartitionedStructure:
    """What `_DirectGCNFused` needs from a graph structure, for a symmetric shared pattern whose rows
    are partitioned (reference-built graphs: grouped-by-target == grouped-by-source == the row CSR)."""

    partitioned = True
    shared = True

    def __init__(self, local: _Csr, n: int, group=None, halo: Optional["HaloExchange"] = None):
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.n = n
        self.lo, self.hi, self.per = row_range(n, self.rank, self.world)
        if local.rowptr.numel() != self.per + 1 or len(local.vals) != 3:
            raise ValueError("PartitionedStructure wants the padded block (per + 1 row pointers) with three value arrays")
        self.local = local
        self.by_dst = self.by_src = [local]
        self.nnz_total = 3 * int(local.col.numel())
        self.halo = self.local_ext = None
        if EXCHANGE_MODE == "halo":
            self.halo = halo if halo is not None else HaloExchange(local.col, n, self.group)   # an existing plan of the SAME block is shared
            self.local_ext = _Csr(local.rowptr, self.halo.col_ext, local.vals)

    def _check(self, t: torch.Tensor):
        if t.shape[0] != self.per:
            raise ValueError(f"row-partitioned layers work on the padded block: expected {self.per} rows, got {t.shape[0]}")

    def fanout_begin(self, x: torch.Tensor, f: int, scales=None, scale_stride: int = 1):
        """Post the exchange of a fan-out now, launch its SpMM later (`fanout_finish`): what is enqueued in between overlaps
        with the NVLink transfer.  None when this structure exchanges by all-gather (nothing to split)."""
        self._check(x)
        if self.halo is None:
            return None
        return _halo_fanout_begin(self.local_ext, self.halo, x, f, scales, scale_stride)

    def fanout_finish(self, pending) -> torch.Tensor:
        return _halo_fanout_finish(pending)

    def fanout(self, x: torch.Tensor, f: int, scales=None, scale_stride: int = 1) -> torch.Tensor:
        self._check(x)
        if self.halo is not None:
            return _halo_fanout(self.local_ext, self.halo, x, f, scales, scale_stride)
        x_full = _all_gather_rows(x.contiguous(), self.group)
        c, per = self.local, self.per
        z = torch.empty((per, 3 * f), dtype=torch.float32, device=x.device)
        if scales is None:
            nat.call("pg_spmm_fanout", nat.ptr(c.rowptr), nat.ptr(c.col), nat.ptr(c.vals[0]), nat.ptr(c.vals[1]), nat.ptr(c.vals[2]), 3, per, f,
                     nat.ptr(x_full), x_full.stride(0), nat.ptr(z), z.stride(0), 0, c.plan(3 * f), nat.stream_ptr())
            return z
        if scale_stride == 1:   # per-node gates: the kernel scales by the SOURCE row, so it needs everybody's
            mine = torch.stack([s.reshape(-1) for s in scales]).contiguous()                 # [3, per]
            allg = torch.empty((self.world * 3, per), dtype=mine.dtype, device=mine.device)
            dist.all_gather_into_tensor(allg, mine, group=self.group)
            full = allg.view(self.world, 3, per).permute(1, 0, 2).reshape(3, self.world * per).contiguous()
            s0, s1, s2 = full[0], full[1], full[2]
        else:
            s0, s1, s2 = scales
        nat.call("pg_spmm_fanout_scaled", nat.ptr(c.rowptr), nat.ptr(c.col), nat.ptr(c.vals[0]), nat.ptr(c.vals[1]), nat.ptr(c.vals[2]), 3, per, f,
                 nat.ptr(x_full), x_full.stride(0), nat.ptr(z), z.stride(0), 0, nat.ptr(s0), nat.ptr(s1), nat.ptr(s2), int(scale_stride),
                 c.plan(3 * f), nat.stream_ptr())
        return z

    def fanin(self, dz: torch.Tensor, f: int, init: Optional[torch.Tensor]) -> torch.Tensor:
        self._check(dz)
        if self.halo is not None:
            return _halo_fanin(self.local_ext, self.halo, dz, f, init)
        return _fanin_exchanged(self.local, dz.contiguous(), self.per, f, init, self.group)


def partitioned_data(x_local: torch.Tensor, local: _Csr, n: int, group=None, halo: Optional[HaloExchange] = None, **extra):
    """Data object for `ProtGramDirectGCN` on this rank's row block.  x_local: [hi - lo, F] (padded here to
    `per` rows); the model must be built with `num_graph_nodes = per` (`row_range(n, rank, world)[2]`).
    Outputs have `per` rows; rows past hi - lo are padding (mask them in the loss)."""
    from .protgram_directgcn import Data, register_structure
    st = PartitionedStructure(local, n, group, halo=halo)
    x = x_local
    if x.shape[0] != st.per:
        x = torch.cat([x, torch.zeros((st.per - x.shape[0], x.shape[1]), dtype=x.dtype, device=x.device)], dim=0)
    # the reference API hands the layer edge tensors; here they only key the structure: this rank's pattern with LOCAL row numbers
    rows = torch.repeat_interleave(torch.arange(st.per, device=local.col.device), local.rowptr[1:] - local.rowptr[:-1])
    ei = torch.stack([local.col.to(torch.int64), rows])      # (source = global column, target = local row)
    ews = (local.vals[0], local.vals[1], local.vals[2])
    register_structure((ei, ei, ei), ews, st.per, st)
    return Data(x=x, edge_index_in=ei, edge_weight_in=ews[0], edge_index_out=ei, edge_weight_out=ews[1],
                edge_index_undirected_norm=ei, edge_weight_undirected_norm=ews[2], num_nodes=st.per, **extra)


PER_NODE_PARAMETERS = ("constant", "C_in_vec", "C_out_vec", "C_directed_vec", "C_undirected_vec", "C_all_vec")


def allreduce_replicated_grads(model: torch.nn.Module, group=None) -> None:
    """Sum the gradients of the replicated parameters (weights, biases, scalar gates, PE table, decoder) over the ranks --
    each rank's backward only saw its rows.  The per-node parameters hold this rank's rows and stay local."""
    group = group if group is not None else dist.group.WORLD
    for name, p in model.named_parameters():
        if not p.requires_grad or name.rsplit(".", 1)[-1] in PER_NODE_PARAMETERS:
            continue
        if p.grad is None:
            # a parameter the loss did not reach on THIS rank (e.g. an all-padding block) still takes part: every rank must issue
            # the same collectives in the same order
            p.grad = torch.zeros_like(p)
        dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=group)
