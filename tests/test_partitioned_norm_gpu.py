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
BACKEND = "nccl"


@pytest.fixture(scope="module")
def world1():
    """ONE one-rank process group for all tests of this module that drive the collective code paths (initialised once per
    process: NCCL communicators are not re-created between tests)."""
    import torch.distributed as dist
    from tests.test_multirank_gloo import _free_port
    created = not dist.is_initialized()
    if created:
        kw = {"device_id": torch.device("cuda", torch.cuda.current_device())} if BACKEND == "nccl" else {}
        dist.init_process_group(BACKEND, init_method=f"tcp://127.0.0.1:{_free_port()}", rank=0, world_size=1, **kw)
    try:
        yield dist.group.WORLD
    finally:
        if created:
            dist.destroy_process_group()


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


def test_normalize_row_partitioned_world1_nccl_equals_single_gpu(world1):
    """The whole driver (owner partition with pg_sort_pairs, all_to_all_single / all_gather over NCCL,
    padding) on a one-rank NCCL group: identical to device_normalize."""
    from protgram_directgcn_b200.host.partitioned import RowPartitionedPropagation, local_csr, normalize_row_partitioned
    n = 1003
    src, dst, cnt = random_count_graph(n, seed=21, density=0.01)
    s, d, w = (torch.from_numpy(a).to(DEV) for a in (src, dst, cnt.astype(np.float32)))
    full = graph_utils.device_normalize(s, d, w, n, 1e-9)
    perm = torch.randperm(s.numel(), generator=torch.Generator().manual_seed(1)).to(DEV)
    res = normalize_row_partitioned(s[perm], d[perm], w[perm], n, group=world1)
    for k in ("rowptr", "col", "val_out", "val_in", "val_und", "in_src", "in_dst", "in_w"):
        assert torch.equal(res[k], full[k]), k
    x = torch.randn(n, 16, device=DEV)
    z = RowPartitionedPropagation.from_local(local_csr(res), n, group=world1)(x)
    ref = torch.zeros(n, 16, device=DEV, dtype=torch.float64)
    rows = torch.repeat_interleave(torch.arange(n, device=DEV), full["rowptr"][1:] - full["rowptr"][:-1])
    ref.index_add_(0, rows, full["val_in"].double().view(-1, 1) * x.double()[full["col"].long()])
    assert float((z[:, :16].double() - ref).abs().max()) <= 1e-5 * float(ref.abs().max())


# ------------------------------------------------------------------ key-range extraction (tables merged by reduce-scatter)
def _synth_corpus(nseq, seq_len, seed=42):
    buf = torch.empty(nseq * (seq_len + 2) + 1, dtype=torch.uint8, device=DEV)
    nat.call("pg_synth_corpus", nat.ptr(buf), 0, nseq, seq_len, seed, 1, nat.stream_ptr())
    return buf


@pytest.mark.parametrize("n,world", [(1, 1), (1, 4), (2, 3), (3, 2), (3, 8)])
def test_key_range_extraction_equals_full_extraction(n, world):
    """csrc/extract_range.cu rank by rank on one GPU (presence MAX-reduced by hand): node codes and the concatenated edge
    lists are exactly what pg_graph_extract_* gives on the whole table.  Few, short sequences so that many n-grams are
    absent (ids != codes) and some key ranges are empty."""
    from protgram_directgcn_b200.host import corpus, data_builder
    buf = _synth_corpus(40 if n == 3 else 12, 30)
    symbols, d_rank = corpus.discover_alphabet(buf)
    sigma = int(symbols.size)
    bins, short = data_builder.count_level(buf, n, d_rank, sigma)
    node_code, src, dst, cnt = data_builder.extract_level(bins, short, n, sigma)
    pow_n = sigma ** n
    assert 0 < node_code.numel() and (n == 1 or node_code.numel() < pow_n)
    codes_per = (pow_n + world - 1) // world
    chunk = codes_per * sigma
    padded = torch.zeros(world * chunk, dtype=torch.int64, device=DEV)
    padded[: bins.numel()] = bins
    st = nat.stream_ptr()
    marks = []
    for r in range(world):
        local = padded[r * chunk:(r + 1) * chunk].clone()
        present = short.clone()
        ws = torch.empty(nat.query("pg_graph_extract_range_ws_bytes", sigma, codes_per), dtype=torch.uint8, device=DEV)
        sizes = torch.zeros(1, dtype=torch.int64, device=DEV)
        nat.call("pg_graph_extract_range_mark", nat.ptr(local), n, sigma, r * codes_per, codes_per, nat.ptr(present), nat.ptr(sizes),
                 nat.ptr(ws), ws.numel(), st)
        marks.append((local, present, ws, int(sizes.item())))
    present = torch.stack([m[1] for m in marks]).max(dim=0).values.contiguous()          # the all-reduce(MAX)
    node_id = torch.empty(pow_n, dtype=torch.int64, device=DEV)
    ws_ids = torch.empty(nat.query("pg_node_ids_ws_bytes", pow_n), dtype=torch.uint8, device=DEV)
    sizes = torch.zeros(1, dtype=torch.int64, device=DEV)
    nat.call("pg_node_ids_from_presence", nat.ptr(present), pow_n, nat.ptr(node_id), nat.ptr(sizes), nat.ptr(ws_ids), ws_ids.numel(), st)
    assert int(sizes.item()) == node_code.numel()
    codes = torch.empty_like(node_code)
    nat.call("pg_node_codes_emit", nat.ptr(present), nat.ptr(node_id), pow_n, nat.ptr(codes), st)
    assert torch.equal(codes, node_code)
    parts = []
    for r, (local, _, ws, e_local) in enumerate(marks):
        s, d, c = (torch.empty(e_local, dtype=torch.int64, device=DEV) for _ in range(3))
        nat.call("pg_graph_extract_range_fill", nat.ptr(local), n, sigma, r * codes_per, codes_per, nat.ptr(node_id), e_local, nat.ptr(s),
                 nat.ptr(d), nat.ptr(c), nat.ptr(ws), ws.numel(), st)
        parts.append((s, d, c))
    assert sum(m[3] for m in marks) == src.numel()
    for i, ref in enumerate((src, dst, cnt)):
        assert torch.equal(torch.cat([p[i] for p in parts]), ref)


def test_fully_partitioned_build_world1_nccl_equals_build_level_graph(world1):
    """build_level_graph_partitioned (reduce_scatter_tensor, key-range extraction, re-deal by source, partitioned
    normalisation) over a one-rank NCCL group == build_level_graph, bitwise, n = 1..3."""
    from protgram_directgcn_b200.host import corpus, data_builder
    buf = _synth_corpus(3000, 60)
    symbols, d_rank = corpus.discover_alphabet(buf)
    for n in (1, 2, 3):
        ref = data_builder.build_level_graph(buf, n, symbols, d_rank, 1e-9)
        got = data_builder.build_level_graph_partitioned(buf, n, symbols, d_rank, 1e-9, world1)
        assert got.node_sequences == ref.node_sequences and got.number_of_edges == ref.number_of_edges
        side = ref._pg_device
        assert torch.equal(got.block["rowptr"], side["rowptr"]) and torch.equal(got.block["col"], side["col"])
        for k in ("val_in", "val_out", "val_und"):
            assert torch.equal(got.block[k], side[k]), (n, k)
        a_out = ref.A_out_w
        assert torch.equal(torch.stack([got.a_out[0], got.a_out[1]]).cpu(), a_out.indices())
        assert torch.equal(got.a_out[2].cpu(), a_out.values())
        z = got.propagation()(torch.ones(got.number_of_nodes, 4, device=DEV))
        assert z.shape == (got.number_of_nodes, 12) and bool(torch.isfinite(z).all())


@pytest.mark.parametrize("n,dims", [(301, [12, 16, 8]), (6000, [64, 128, 128])])
def test_row_partitioned_model_world1_nccl_equals_plain_model(n, dims, world1):
    """ProtGramDirectGCN on the padded row block of a one-rank NCCL group == the plain model on the same graph (forward,
    loss, every gradient); the second shape takes the tensor-core path (scaled fan-out with exchanged gates in backward)."""
    import protgram_directgcn_b200 as pg
    from protgram_directgcn_b200.host import partitioned as part
    src, dst, cnt = random_count_graph(n, seed=5, density=min(0.05, 20.0 / n))
    s, d, w = (torch.from_numpy(a).to(DEV) for a in (src, dst, cnt.astype(np.float32)))
    full = graph_utils.device_normalize(s, d, w, n, 1e-9)
    ei = graph_utils.csr_to_coo_indices(full["rowptr"], full["col"], n).flip(0).contiguous()
    x = torch.randn(n, dims[0], device=DEV)
    y = torch.randint(0, 4, (n,), device=DEV)
    torch.manual_seed(0)
    plain = pg.ProtGramDirectGCN(dims, n, 4, 1, 0, 0, 0.0, True).to(DEV)
    twin = pg.ProtGramDirectGCN(dims, n, 4, 1, 0, 0, 0.0, True).to(DEV)
    twin.load_state_dict(plain.state_dict())
    plain.eval(), twin.eval()
    data = pg.Data(x=x, edge_index_in=ei, edge_weight_in=full["val_in"], edge_index_out=ei, edge_weight_out=full["val_out"],
                   edge_index_undirected_norm=ei, edge_weight_undirected_norm=full["val_und"])
    logp, emb = plain(data)
    torch.nn.functional.nll_loss(logp, y).backward()
    res = part.normalize_row_partitioned(s, d, w, n, group=world1)
    pdata = part.partitioned_data(x, part.local_csr(res), n, group=world1)
    logp2, emb2 = twin(pdata)
    torch.nn.functional.nll_loss(logp2, y).backward()
    part.allreduce_replicated_grads(twin, group=world1)
    assert float((logp2 - logp).abs().max()) <= 2e-5 and float((emb2 - emb).abs().max()) <= 2e-5
    for (k, p), (_, p2) in zip(plain.named_parameters(), twin.named_parameters()):
        if p.grad is not None:
            assert float((p2.grad - p.grad).abs().max()) <= 1e-4 * max(1.0, float(p.grad.abs().max())), k


def test_loaded_graph_with_csr_sidecar_feeds_the_layer(tmp_path):
    import protgram_directgcn_b200 as pg
    from tests.helpers import golden_edges, load
    g = load("build_protein")
    src, dst, w = golden_edges(g, 3)
    graph = pg.DirectedNgramGraph.from_edge_arrays({i: s for i, s in enumerate(g["n3_nodes"])}, src, dst, w.astype(np.float32),
                                                   n_value=3, assume_coalesced=True)
    path = str(tmp_path / "g.pkl")
    pg.DataUtils.save_object(graph, path)
    loaded = pg.DataUtils.load_object(path)
    assert loaded.__dict__.get("_pg_sidecar") is not None
    torch.manual_seed(0)
    model = pg.ProtGramDirectGCN([16, 32, 8], graph.number_of_nodes, 3, 3, 0, 0, 0.0, True).to(DEV).eval()
    x = torch.randn(graph.number_of_nodes, 16, device=DEV)
    assert torch.equal(model(loaded.gcn_data(x, DEV))[1], model(graph.gcn_data(x, DEV))[1])


def test_partitioned_halo_exchange_equals_allgather_exchange_bitwise(world1, monkeypatch):
    """One-rank NCCL group: the halo path (split operand, feature-column chunks, scaled fan-out with exchanged gates, fan-in)
    against the round-1 all-gather path and against the plain single-GPU kernels -- bitwise (same CSR order per row)."""
    from protgram_directgcn_b200.host import partitioned as part
    n, f = 5003, 64
    src, dst, cnt = random_count_graph(n, seed=31, density=0.004)
    s, d, w = (torch.from_numpy(a).to(DEV) for a in (src, dst, cnt.astype(np.float32)))
    full = graph_utils.device_normalize(s, d, w, n, 1e-9)
    csr = part.slice_rows(full["rowptr"], full["col"], [full["val_in"], full["val_out"], full["val_und"]], 0, n, n)
    x = torch.randn(n, f, device=DEV)
    dz = torch.randn(n, 3 * f, device=DEV)
    init = torch.randn(n, f, device=DEV)
    gates = tuple(torch.rand(n, device=DEV) + 0.5 for _ in range(3))
    outs = {}
    for mode, chunks in (("allgather", 1), ("halo", 1), ("halo", 4)):
        monkeypatch.setattr(part, "EXCHANGE_MODE", mode)
        monkeypatch.setattr(part, "PIPELINE_CHUNKS", chunks)
        monkeypatch.setattr(part, "PIPELINE_MIN_BYTES", 0)
        st = part.PartitionedStructure(csr, n, world1)
        outs[(mode, chunks)] = (st.fanout(x, f), st.fanout(x, f, scales=gates, scale_stride=1),
                                st.fanout(x, f, scales=tuple(g[:1] for g in gates), scale_stride=0), st.fanin(dz, f, init))
    ref = outs[("allgather", 1)]
    for key, got in outs.items():
        for a, b in zip(got, ref):
            assert torch.equal(a, b), key


def _rmat_edges(log2_nodes, edges_per_node, seed=42):
    """bench.py's R-MAT generator (a, b, c, d = .57, .19, .19, .05; ids randomly relabelled; integer weights)."""
    n, e = 1 << log2_nodes, (1 << log2_nodes) * edges_per_node
    g = torch.Generator(device=DEV).manual_seed(seed)
    src = torch.zeros(e, dtype=torch.int64, device=DEV)
    dst = torch.zeros(e, dtype=torch.int64, device=DEV)
    for _ in range(log2_nodes):
        r = torch.rand(e, generator=g, device=DEV)
        src = src * 2 + (r >= 0.76).to(torch.int64)
        dst = dst * 2 + (((r >= 0.57) & (r < 0.76)) | (r >= 0.95)).to(torch.int64)
    w = torch.randint(1, 8, (e,), generator=g, device=DEV).to(torch.float32)
    perm = torch.randperm(n, generator=g, device=DEV)
    return n, perm[src], perm[dst], w


@pytest.mark.parametrize("world,chunks", [(4, 1), (8, 2)])
def test_partitioned_spmm_on_rmat_blocks_vs_oracle_sampled_rows(world, chunks):
    """Config C5's shape at test size (VERDICT r1 missing #1): an R-MAT power-law digraph (2^17 nodes, 2 M edges, hub rows of
    several thousand entries) through the reference normalisation, its rows cut into `world` blocks that are played one after
    the other on this GPU with the PRODUCT's halo plan (`partitioned.halo_plan`: referenced remote rows, renumbered columns),
    pack kernel and split-operand kernels.  Each block's forward fan-out, gated (backward) fan-out and fan-in must
      (1) equal the same rows of the single-GPU kernels bit for bit, and
      (2) agree with oracle/directgcn_oracle.py's `propagate` (reference message passing, torch CPU) on 256 sampled rows
          to 1e-4 relative (the north star's bar; fp32 sums over hub rows differ by summation order only)."""
    from oracle import directgcn_oracle
    from protgram_directgcn_b200.host import partitioned as part
    from protgram_directgcn_b200.host.protgram_directgcn import _Csr
    n, src, dst, w = _rmat_edges(17, 16)
    s, d, wv = graph_utils.device_coalesce(src, dst, w, n)
    full = graph_utils.device_normalize(s, d, wv, n, 1e-9)
    rowptr, col = full["rowptr"], full["col"]
    vals = [full["val_in"], full["val_out"], full["val_und"]]
    F = 64
    torch.manual_seed(1)
    x = torch.randn(n, F, device=DEV)
    dz = torch.randn(n, 3 * F, device=DEV)
    gates = torch.rand(n, 4, device=DEV) + 0.5
    st = nat.stream_ptr()
    whole = _Csr(rowptr, col, vals)
    plan = whole.plan(3 * F)
    assert whole._plan.n_long > 0                                       # hub rows take the long-row split
    vp = [nat.ptr(v) for v in vals]
    z_ref, zs_ref, y_ref = torch.empty(n, 3 * F, device=DEV), torch.empty(n, 3 * F, device=DEV), torch.empty(n, F, device=DEV)
    g_cols = [gates[:, k].contiguous() for k in range(3)]
    nat.call("pg_spmm_fanout", nat.ptr(rowptr), nat.ptr(col), *vp, 3, n, F, nat.ptr(x), F, nat.ptr(z_ref), 3 * F, 0, plan, st)
    nat.call("pg_spmm_fanout_scaled", nat.ptr(rowptr), nat.ptr(col), *vp, 3, n, F, nat.ptr(x), F, nat.ptr(zs_ref), 3 * F, 0,
             nat.ptr(g_cols[0]), nat.ptr(g_cols[1]), nat.ptr(g_cols[2]), 1, plan, st)
    nat.call("pg_spmm_fanin", nat.ptr(rowptr), nat.ptr(col), *vp, 3, n, F, nat.ptr(dz), 3 * F, 0, None, 0, nat.ptr(y_ref), F, 0, plan, st)
    halo_rows = []
    z_all, zs_all, y_all = (torch.empty_like(t) for t in (z_ref, zs_ref, y_ref))
    w_chunk = F // chunks
    for r in range(world):
        lo, hi, per = row_range(n, r, world)
        blk = part.slice_rows(rowptr, col, vals, lo, hi, per)
        need, need_counts, col_ext = part.halo_plan(blk.col, lo, per, world)
        assert int(need_counts[r]) == 0 and int(need_counts.sum()) == need.numel()
        halo_rows.append(int(need.numel()))
        bp = [nat.ptr(v) for v in blk.vals]
        bplan = blk.plan(3 * F)
        # the peers' pack step: owner o gathers the rows this block asked it for (local row numbers), in `need` order
        def halo_of(t, c0, cw):
            out = torch.empty((need.numel(), cw), dtype=torch.float32, device=DEV)
            pos = 0
            for o in range(world):
                k = int(need_counts[o])
                if k:
                    olo = row_range(n, o, world)[0]
                    idx = (need[pos:pos + k] - olo).contiguous()
                    own_rows = t[olo:olo + per]
                    nat.call("pg_gather_rows", nat.ptr(own_rows[:, c0:]), t.stride(0), nat.ptr(idx), k, cw, nat.ptr(out[pos:]), cw, st)
                pos += k
            return out
        xl, gl, dzl = x[lo:lo + per], gates[lo:lo + per], dz[lo:lo + per]
        z, zs = torch.empty(per, 3 * F, device=DEV), torch.empty(per, 3 * F, device=DEV)
        g_ext = torch.cat([gl, halo_of(gates, 0, 4)]).reshape(-1)
        for c0 in range(0, F, w_chunk):
            hx = halo_of(x, c0, w_chunk)
            for out, sp, stride in ((z, (None, None, None), 0), (zs, tuple(nat.ptr(g_ext[k:]) for k in range(3)), 4)):
                nat.call("pg_spmm_fanout_split", nat.ptr(blk.rowptr), nat.ptr(col_ext), *bp, 3, hi - lo, w_chunk,
                         nat.spmm_operand(xl[:, c0:c0 + w_chunk], hx, per), nat.ptr(out[:, c0:]), 3 * F, 0, F, sp[0], sp[1], sp[2], stride, bplan, st)
        y = torch.empty(per, F, device=DEV)
        nat.call("pg_spmm_fanin_split", nat.ptr(blk.rowptr), nat.ptr(col_ext), *bp, 3, hi - lo, F, nat.spmm_operand(dzl, halo_of(dz, 0, 3 * F), per),
                 0, F, None, 0, nat.ptr(y), F, 0, bplan, st)
        z_all[lo:hi], zs_all[lo:hi], y_all[lo:hi] = z[: hi - lo], zs[: hi - lo], y[: hi - lo]
    assert torch.equal(z_all, z_ref) and torch.equal(zs_all, zs_ref) and torch.equal(y_all, y_ref)
    # power-law graph, random relabelling: a block references well under all remote rows (what makes the halo exchange pay)
    assert max(halo_rows) < 0.8 * (n - n // world), halo_rows
    # (2) reference message passing on 256 sampled target rows
    sample = torch.from_numpy(np.random.default_rng(0).choice(n, 256, replace=False))
    rows_of = torch.repeat_interleave(torch.arange(n, device=DEV), rowptr[1:] - rowptr[:-1]).cpu()
    keep = torch.zeros(n, dtype=torch.bool)
    keep[sample] = True
    keep = keep[rows_of]
    ei = torch.stack([col.cpu().long()[keep], rows_of[keep]])          # (source = column, target = row)
    xc, dzc, gc = x.cpu(), dz.cpu(), gates.cpu()
    err = lambda a, b: float((a - b).abs().max() / b.abs().max())
    for v in range(3):
        ew = vals[v].cpu()[keep]
        ref = directgcn_oracle.propagate(ei, xc, ew)[sample]
        assert err(z_all[sample, v * F:(v + 1) * F].cpu(), ref) <= 1e-4, v
        ref_s = directgcn_oracle.propagate(ei, xc * gc[:, v:v + 1], ew)[sample]
        assert err(zs_all[sample, v * F:(v + 1) * F].cpu(), ref_s) <= 1e-4, v
    ref_y = sum(directgcn_oracle.propagate(ei, dzc[:, v * F:(v + 1) * F], vals[v].cpu()[keep]) for v in range(3))[sample]
    assert err(y_all[sample].cpu(), ref_y) <= 1e-4


def test_halo_push_and_wait_kernels_emulated_ranks():
    """csrc/peer.cu on ONE GPU: three ranks' buffers live in this process (plain pointers instead of CUDA-IPC mappings; the IPC
    plumbing itself runs in tools/multigpu_check.py on real peers).  Every rank pushes the rows each other rank asked for into
    that rank's receive slot (work in rotated peer order) and publishes the epoch; the wait kernels return at once (the pushes
    are earlier in stream order); a wait for an epoch nobody publishes gives up after its bounded spin and raises the error
    word instead of hanging."""
    import ctypes
    per, w, world = 1000, 64, 3
    g = torch.Generator(device=DEV).manual_seed(0)
    x = [torch.randn(per, w, device=DEV, generator=g) for _ in range(world)]
    counts = [[0, 700, 13], [300, 0, 0], [129, 511, 0]]                 # counts[r][p]: rows rank r sends to rank p
    serve = [torch.randint(0, per, (sum(counts[r]),), device=DEV, generator=g) for r in range(world)]
    off_on = [[sum(counts[q][p] for q in range(r)) for p in range(world)] for r in range(world)]    # rank r's first row in p's halo order
    recv = [torch.full((sum(counts[q][p] for q in range(world)), w), float("nan"), device=DEV) for p in range(world)]
    flags = [torch.zeros(64, dtype=torch.int32, device=DEV) for _ in range(world)]
    done = torch.zeros(1, dtype=torch.int32, device=DEV)
    err = torch.zeros(1, dtype=torch.int32, device=DEV)
    st = nat.stream_ptr()
    P = ctypes.c_void_p
    for epoch in (1, 2):
        for r in range(world):
            begin = [0]
            for p in range(world):
                begin.append(begin[-1] + counts[r][p])
            rb = (ctypes.c_int64 * (world + 1))(*begin)
            dst = (P * world)(*[P(recv[p].data_ptr() + off_on[r][p] * w * 4) if counts[r][p] else P(None) for p in range(world)])
            flg = (P * world)(*[P(flags[p].data_ptr() + 4 * r) if p != r else P(None) for p in range(world)])
            nat.call("pg_halo_push", nat.ptr(x[r]), w, nat.ptr(serve[r]), rb, dst, flg, world, r, w, w, epoch, nat.ptr(done), st)
        for r in range(world):
            nat.call("pg_halo_wait", nat.ptr(flags[r]), world, r, epoch, nat.ptr(err), st)
        torch.cuda.synchronize()
        for p in range(world):
            want = torch.cat([x[r][serve[r][sum(counts[r][:p]):sum(counts[r][:p + 1])]] for r in range(world)])
            assert torch.equal(recv[p], want)
            assert all(int(flags[p][r]) == epoch for r in range(world) if r != p)
        assert int(err) == 0 and int(done) == 0
        x = [t + 1.0 for t in x]
    import time
    t0 = time.time()
    nat.call("pg_halo_wait", nat.ptr(flags[0]), world, 0, 7, nat.ptr(err), st)     # nobody publishes epoch 7
    torch.cuda.synchronize()
    assert int(err) in (2, 3) and time.time() - t0 < 30.0                          # 1 + a peer that stayed silent
