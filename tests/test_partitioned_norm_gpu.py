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


def test_partitioned_fanin_column_chunks_equal_single_exchange(world1):
    from protgram_directgcn_b200.host import partitioned as part
    n, f = 5003, 64
    src, dst, cnt = random_count_graph(n, seed=31, density=0.004)
    s, d, w = (torch.from_numpy(a).to(DEV) for a in (src, dst, cnt.astype(np.float32)))
    full = graph_utils.device_normalize(s, d, w, n, 1e-9)
    csr = part.slice_rows(full["rowptr"], full["col"], [full["val_in"], full["val_out"], full["val_und"]], 0, n, n)
    dz = torch.randn(n, 3 * f, device=DEV)
    init = torch.randn(n, f, device=DEV)
    one = part._fanin_exchanged(csr, dz, n, f, init, world1)
    for limit in (n * 3 * f * 4 // 2, n * 3 * f * 4 // 7):
        assert torch.equal(part._fanin_exchanged(csr, dz, n, f, init, world1, limit=limit), one)
