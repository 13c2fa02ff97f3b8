"""Row-partitioned normalisation on the GPU: the row-block kernels of
csrc/graph_rows.cu against (1) the single-GPU kernels on the whole graph -- BITWISE, block by
block -- and (2) the full-graph CPU oracle.  The ranks are played one after the other on one GPU
(the phases of RowBlockNormalizer are local; the per-node vectors that cross ranks are concatenated
here instead of all-gathered; the collective plumbing itself is covered over gloo in
tests/test_multirank_gloo.py and over NCCL by tools/multigpu_check.py)."""
import numpy as np
import pytest
import torch

from protgram_directgcn_b200 import _native as nat
from protgram_directgcn_b200.host import graph_utils
from protgram_directgcn_b200.host.partitioned import RowBlockNormalizer, row_range
from tests.test_multirank_gloo import check_blocks_against_oracle, random_count_graph

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _play_ranks(src, dst, w, n, world, seed=0, eps=1e-9):
    """-> list of per-rank result dicts (device tensors), rowptr NOT padded."""
    g = torch.Generator().manual_seed(seed)
    s, d, wt = (torch.from_numpy(a).to(DEV) for a in (src, dst, w))
    blocks = []
    for r in range(world):
        lo, hi, per = row_range(n, r, world)
        mine = torch.nonzero((s >= lo) & (s < hi)).flatten()
        theirs = torch.nonzero((d >= lo) & (d < hi)).flatten()
        mine = mine[torch.randperm(mine.numel(), generator=g).to(DEV)]          # any order must do
        theirs = theirs[torch.randperm(theirs.numel(), generator=g).to(DEV)]
        blocks.append(RowBlockNormalizer(s[mine], d[mine], wt[mine], s[theirs], d[theirs], wt[theirs], n, lo, hi, eps))
    sums = [b.degree_sums() for b in blocks]
    rs_out = torch.cat([x[0] for x in sums])
    rs_in = torch.cat([x[1] for x in sums])
    deg = torch.cat([b.structure() for b in blocks])
    assert rs_out.numel() == n and deg.numel() == n
    return [b.values(rs_out, rs_in, deg) for b in blocks], (rs_out, rs_in, deg)


@pytest.mark.parametrize("n,world,density", [(5, 1, 0.3), (5, 4, 0.3), (203, 2, 0.05), (203, 3, 0.05), (4001, 8, 0.004)])
def test_row_blocks_bitwise_equal_single_gpu_and_match_oracle(n, world, density):
    src, dst, cnt = random_count_graph(n, seed=n + world, density=density, isolated=3 if n > 50 else 1)
    w = cnt.astype(np.float32)
    full = graph_utils.device_normalize(torch.from_numpy(src).to(DEV), torch.from_numpy(dst).to(DEV), torch.from_numpy(w).to(DEV), n, 1e-9)
    blocks, (rs_out, rs_in, deg) = _play_ranks(src, dst, w, n, world)
    rp_full = full["rowptr"].cpu().numpy()
    native = np.zeros(n, dtype=np.int64)
    native[src[src == dst]] = 1
    assert np.array_equal(deg.cpu().numpy(), np.diff(rp_full) + native)
    assert np.array_equal(rs_out.cpu().numpy(), np.bincount(src, weights=cnt, minlength=n))
    assert np.array_equal(rs_in.cpu().numpy(), np.bincount(dst, weights=cnt, minlength=n))
    in_rows_full = full["in_src"].cpu().numpy()
    host = {}
    for r, b in enumerate(blocks):
        lo, hi, per = row_range(n, r, world)
        p0, p1 = int(rp_full[lo]), int(rp_full[hi])
        assert b["pattern_nnz"] == p1 - p0
        if hi > lo:
            assert np.array_equal(b["rowptr"].cpu().numpy(), rp_full[lo:hi + 1] - p0)
        assert torch.equal(b["col"], full["col"][p0:p1])
        for k in ("val_out", "val_in", "val_und"):
            assert torch.equal(b[k], full[k][p0:p1]), (k, r)             # bitwise
        e0, e1 = np.searchsorted(in_rows_full, [lo, hi])
        for k in ("in_src", "in_dst", "in_w"):
            assert torch.equal(b[k], full[k][e0:e1]), (k, r)
        rp = np.full(per + 1, p1 - p0, dtype=np.int64)
        rp[: hi - lo + 1] = b["rowptr"].cpu().numpy()
        host[r] = {k: v.cpu().numpy() for k, v in b.items() if torch.is_tensor(v)}
        host[r]["rowptr"] = rp
    check_blocks_against_oracle(host, src, dst, cnt, n, world)


def test_row_blocks_float_weights_and_eps():
    """Non-integer weights (benchmarker-style tables): fp64 degree atomics may differ in the last bit
    between the two summation orders, so values are held to 1e-6 instead of bitwise."""
    n, world = 301, 3
    src, dst, cnt = random_count_graph(n, seed=9, density=0.04)
    w = (np.random.default_rng(2).random(src.size) * 3 + 0.01).astype(np.float32)
    full = graph_utils.device_normalize(torch.from_numpy(src).to(DEV), torch.from_numpy(dst).to(DEV), torch.from_numpy(w).to(DEV), n, 1e-5)
    blocks, _ = _play_ranks(src, dst, w, n, world, eps=1e-5)
    for k in ("val_out", "val_in", "val_und"):
        got = torch.cat([b[k] for b in blocks])
        assert torch.equal(torch.cat([b["col"] for b in blocks]), full["col"])
        assert float(((got - full[k]).abs() / full[k].abs()).max()) <= 1e-6, k


def test_row_block_rejects_foreign_edges():
    n = 50
    src, dst, cnt = random_count_graph(n, seed=3, density=0.1, isolated=1)
    s, d, w = (torch.from_numpy(a).to(DEV) for a in (src, dst, cnt.astype(np.float32)))
    blk = RowBlockNormalizer(s, d, w, s[:0], d[:0], w[:0], n, 0, 25)       # out-edges of ALL rows handed to block [0, 25)
    with pytest.raises(ValueError):
        blk.structure()
    lib = nat.load()
    assert lib.pg_normalize_rows_sizes(None, None, 0, None, None, 0, 10, 8, 5, None, None, 0, None) == -1   # row_lo + rows > num_nodes


def test_normalize_row_partitioned_world1_nccl_equals_single_gpu():
    """The whole driver (owner partition with pg_sort_pairs, all_to_all_single / all_gather over NCCL,
    padding) on a one-rank NCCL group: identical to device_normalize."""
    import torch.distributed as dist
    from protgram_directgcn_b200.host.partitioned import RowPartitionedPropagation, local_csr, normalize_row_partitioned
    from tests.test_multirank_gloo import _free_port
    n = 1003
    src, dst, cnt = random_count_graph(n, seed=21, density=0.01)
    s, d, w = (torch.from_numpy(a).to(DEV) for a in (src, dst, cnt.astype(np.float32)))
    full = graph_utils.device_normalize(s, d, w, n, 1e-9)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{_free_port()}", rank=0, world_size=1,
                            device_id=torch.device("cuda", torch.cuda.current_device()))
    try:
        perm = torch.randperm(s.numel(), generator=torch.Generator().manual_seed(1)).to(DEV)
        res = normalize_row_partitioned(s[perm], d[perm], w[perm], n)
        for k in ("rowptr", "col", "val_out", "val_in", "val_und", "in_src", "in_dst", "in_w"):
            assert torch.equal(res[k], full[k]), k
        x = torch.randn(n, 16, device=DEV)
        z = RowPartitionedPropagation.from_local(local_csr(res), n)(x)
        ref = torch.zeros(n, 16, device=DEV, dtype=torch.float64)
        rows = torch.repeat_interleave(torch.arange(n, device=DEV), full["rowptr"][1:] - full["rowptr"][:-1])
        ref.index_add_(0, rows, full["val_in"].double().view(-1, 1) * x.double()[full["col"].long()])
        assert float((z[:, :16].double() - ref).abs().max()) <= 1e-5 * float(ref.abs().max())
    finally:
        dist.destroy_process_group()
