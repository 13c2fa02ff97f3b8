"""N > 1 paths on CPU: world_size-2 gloo process groups, native entry points replaced by their
executable spec (tests/kernel_spec.py) so only the host-side sharding / collective logic is under
test here.  The same code runs over NCCL on the GPU box (bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.helpers import BUILD_FIXTURES, MATS, fasta_sequences, load


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _init(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from protgram_directgcn_b200 import _native as nat
    from tests import kernel_spec
    kernel_spec.install_plain(nat)
    return nat


def _builder_worker(rank, world, port, fasta_path, out_dir, n_max, q, key_range=False):
    try:
        _init(rank, world, port)
        import protgram_directgcn_b200 as pg
        if key_range:    # force the merge policy of tables beyond L2 (reduce-scatter over key ranges, edges gathered on rank 0)
            from protgram_directgcn_b200.host import data_builder
            data_builder.KEY_RANGE_MERGE_MIN_TABLE_BYTES = 0
        cfg = pg.Config()
        cfg.GCN_INPUT_FASTA_PATH = fasta_path
        cfg.BASE_OUTPUT_DIR = out_dir
        cfg.GRAPH_OBJECTS_DIR = os.path.join(out_dir, "graphs")
        cfg.GCN_NGRAM_MAX_N = n_max
        cfg.GRAPH_BUILDER_PROCESS_GROUP = dist.group.WORLD
        pg.GraphBuilder(cfg).run()
        dist.barrier()
        q.put((rank, "ok"))
    except Exception as exc:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,key_range", [("build_protein", False), ("build_ragged", False), ("build_protein", True), ("build_ragged", True)])
def test_sharded_builder_equals_reference(name, key_range, tmp_path):
    """Corpus split over 2 ranks by sequence range, tables merged by all-reduce: bit-exact nodes,
    edges and counts against the reference goldens (== the single-rank result)."""
    g = load(name)
    fasta = fasta_sequences(str(g["fasta"]), tmp_path)
    n_max = BUILD_FIXTURES[name]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_builder_worker, args=(r, 2, port, fasta, str(tmp_path), n_max, q, key_range)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
    import protgram_directgcn_b200 as pg
    from tests.test_host_logic_cpu import check_graph_against_golden
    for n in range(1, n_max + 1):
        graph = pg.DataUtils.load_object(os.path.join(str(tmp_path), "graphs", f"ngram_graph_n{n}.pkl"))
        check_graph_against_golden(graph, g, n)


def _prop_worker(rank, world, port, payload, q):
    try:
        nat = _init(rank, world, port)
        from protgram_directgcn_b200.host.partitioned import RowPartitionedPropagation, row_range
        from protgram_directgcn_b200.host.protgram_directgcn import _Csr
        rowptr, col, vals, x, gz, n, symmetric, t_rowptr_list, t_col_list, t_vals_list = payload
        lo, hi, per = row_range(n, rank, world)
        transposed = None
        if not symmetric:
            transposed = _Csr(t_rowptr_list[rank], t_col_list[rank], t_vals_list[rank])
        prop = RowPartitionedPropagation(rowptr, col, vals, n, symmetric=symmetric, transposed=transposed)
        xl = x[lo:hi].clone().requires_grad_(True)
        z = prop(xl)
        gl = torch.zeros_like(z)
        gl[: hi - lo] = gz[lo:hi]
        z.backward(gl)
        q.put((rank, z.detach()[: hi - lo].numpy(), xl.grad.numpy()))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc(), None))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("symmetric", [True, False])
def test_row_partitioned_propagation_equals_single(symmetric):
    from protgram_directgcn_b200.host.partitioned import row_range
    rng = np.random.default_rng(4 if symmetric else 5)
    n, f, world = 301, 12, 2
    dense = [np.where(rng.random((n, n)) < 0.03, rng.standard_normal((n, n)), 0.0).astype(np.float32) for _ in range(3)]
    mask = (dense[0] != 0) | (dense[1] != 0) | (dense[2] != 0)
    if symmetric:
        mask = mask | mask.T
        dense = [np.where(mask, (d + d.T) / 2 + 0.1, 0).astype(np.float32) for d in dense]
    else:
        dense = [np.where(mask, d + 0.1, 0).astype(np.float32) for d in dense]
    rows, cols = np.nonzero(mask)
    rowptr = torch.from_numpy(np.concatenate([[0], np.cumsum(np.bincount(rows, minlength=n))]).astype(np.int64))
    col = torch.from_numpy(cols.astype(np.int32))
    vals = [torch.from_numpy(d[rows, cols]) for d in dense]
    x = torch.from_numpy(rng.standard_normal((n, f)).astype(np.float32))
    gz = torch.from_numpy(rng.standard_normal((n, 3 * f)).astype(np.float32))
    # single-process truth
    z_ref = torch.cat([torch.from_numpy(d) @ x for d in dense], dim=1)
    dx_ref = sum(torch.from_numpy(d).t() @ gz[:, v * f:(v + 1) * f] for v, d in enumerate(dense))
    # source-grouped CSR of each rank's row block (general case): rows = global source id, cols = local target index
    t_rp, t_col, t_vals = [], [], []
    for r in range(world):
        lo, hi, per = row_range(n, r, world)
        blk = mask[lo:hi]
        tr, ts = np.nonzero(blk)            # local target, global source
        order = np.lexsort((tr, ts))
        ts_s, tr_s = ts[order], tr[order]
        t_rp.append(torch.from_numpy(np.concatenate([[0], np.cumsum(np.bincount(ts_s, minlength=world * per))]).astype(np.int64)))
        t_col.append(torch.from_numpy(tr_s.astype(np.int32)))
        t_vals.append([torch.from_numpy(d[lo:hi][tr_s, ts_s]) for d in dense])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    payload = (rowptr, col, vals, x, gz, n, symmetric, t_rp, t_col, t_vals)
    procs = [ctx.Process(target=_prop_worker, args=(r, world, port, payload, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=180) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    assert all(not isinstance(r[1], str) for r in res), res
    z = np.concatenate([r[1] for r in res])
    dx = np.concatenate([r[2][: row_range(n, r[0], world)[1] - row_range(n, r[0], world)[0]] for r in res])
    assert np.max(np.abs(z - z_ref.numpy())) <= 1e-4 * np.max(np.abs(z_ref.numpy()))
    assert np.max(np.abs(dx - dx_ref.numpy())) <= 1e-4 * np.max(np.abs(dx_ref.numpy()))


def random_count_graph(n, seed, density=0.02, isolated=3):
    """Coalesced directed edge table with integer counts, reciprocal pairs, native self loops and a few
    nodes without any edge (what the n-gram builder can emit)."""
    rng = np.random.default_rng(seed)
    mask = rng.random((n, n)) < density
    mask[np.arange(0, n, 7), np.arange(0, n, 7)] = True        # native self loops
    mask |= (rng.random((n, n)) < density / 2) & mask.T          # extra reciprocal pairs
    dead = rng.choice(n, size=isolated, replace=False)
    mask[dead, :] = False
    mask[:, dead] = False
    src, dst = np.nonzero(mask)
    cnt = rng.integers(1, 400, size=src.size).astype(np.int64)
    return src.astype(np.int64), dst.astype(np.int64), cnt


def check_blocks_against_oracle(blocks, src, dst, cnt, n, world, bitwise_vs=None):
    """blocks: rank -> dict of numpy arrays from normalize_row_partitioned.  Reassemble the full
    pattern / values / A_in_w and hold them against the full-graph oracle."""
    from oracle import graph_oracle
    from protgram_directgcn_b200.host.partitioned import row_range
    mats = graph_oracle.normalise_all(src, dst, cnt, n)
    rows, cols, vals, ain = [], [], {"val_out": [], "val_in": [], "val_und": []}, [[], [], []]
    for r in range(world):
        b = blocks[r]
        lo, hi, per = row_range(n, r, world)
        rp = b["rowptr"]
        assert rp.shape[0] == per + 1 and rp[0] == 0 and rp[hi - lo] == b["col"].shape[0] and np.all(rp[hi - lo:] == rp[hi - lo])
        rows.append(np.repeat(np.arange(lo, hi), np.diff(rp[: hi - lo + 1])))
        cols.append(b["col"].astype(np.int64))
        for k in vals:
            vals[k].append(b[k])
        for i, k in enumerate(("in_src", "in_dst", "in_w")):
            ain[i].append(b[k])
    rows, cols = np.concatenate(rows), np.concatenate(cols)
    for name, key in (("mathcal_A_out", "val_out"), ("mathcal_A_in", "val_in"), ("A_undirected_norm_sparse", "val_und")):
        r_ref, c_ref, v_ref = mats[name]
        assert np.array_equal(rows, r_ref) and np.array_equal(cols, c_ref), name
        got = np.concatenate(vals[key])
        assert np.max(np.abs(got - v_ref) / np.abs(v_ref)) <= 2e-7, name
    r_ref, c_ref, v_ref = mats["A_in_w"]
    assert np.array_equal(np.concatenate(ain[0]), r_ref) and np.array_equal(np.concatenate(ain[1]), c_ref)
    assert np.array_equal(np.concatenate(ain[2]), v_ref)


def _norm_worker(rank, world, port, payload, q):
    try:
        _init(rank, world, port)
        from protgram_directgcn_b200.host.partitioned import (RowPartitionedPropagation, local_csr, normalize_row_partitioned,
                                                              row_range)
        src, dst, cnt, n, x = payload
        lo, hi, per = row_range(n, rank, world)
        mine = (src >= lo) & (src < hi)
        res = normalize_row_partitioned(torch.from_numpy(src[mine]), torch.from_numpy(dst[mine]),
                                        torch.from_numpy(cnt[mine].astype(np.float32)), n)
        if src.size == 0:           # a graph without any edge: empty matrices on every rank, like the reference
            q.put((rank, {"pattern_nnz": res["pattern_nnz"], "rowptr": res["rowptr"].numpy()}))
            return
        prop = RowPartitionedPropagation.from_local(local_csr(res), n)
        z = prop(torch.from_numpy(x[lo:hi]))[: hi - lo]
        out = {k: res[k].numpy() for k in ("rowptr", "col", "val_out", "val_in", "val_und", "in_src", "in_dst", "in_w")}
        out["z"] = z.numpy()
        q.put((rank, out))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,world", [(203, 2), (5, 3), (5, 4)])
def test_row_partitioned_normalisation_equals_oracle(n, world):
    """every rank holds the out-edges of its rows; one all-to-all of edges + three
    all-gathered per-node vectors give each rank its rows of the three propagation matrices and of
    A_in_w.  Reassembled blocks == the full-graph oracle; the blocks then feed the partitioned SpMM.
    (5 nodes: per = 2, so the third rank owns a single row and a fourth rank none at all.)"""
    from oracle import graph_oracle
    src, dst, cnt = random_count_graph(n, seed=n, density=0.05 if n > 50 else 0.3, isolated=3 if n > 50 else 1)
    rng = np.random.default_rng(1)
    x = rng.standard_normal((n, 8)).astype(np.float32)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_norm_worker, args=(r, world, port, (src, dst, cnt, n, x), q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert all(not isinstance(v, str) for v in res.values()), res
    check_blocks_against_oracle(res, src, dst, cnt, n, world)
    mats = graph_oracle.normalise_all(src, dst, cnt, n)
    z = np.concatenate([res[r]["z"] for r in range(world)])
    for v, name in enumerate(("mathcal_A_in", "mathcal_A_out", "A_undirected_norm_sparse")):
        rr, cc, vv = mats[name]
        dense = np.zeros((n, n), dtype=np.float64)
        dense[rr, cc] = vv
        ref = dense @ x.astype(np.float64)
        assert np.max(np.abs(z[:, v * 8:(v + 1) * 8] - ref)) <= 1e-5 * max(1.0, np.max(np.abs(ref))), name


def _partitioned_build_worker(rank, world, port, seqs, n_max, q):
    try:
        nat = _init(rank, world, port)
        from protgram_directgcn_b200.host import corpus, data_builder
        per = (len(seqs) + world - 1) // world
        mine = seqs[rank * per:(rank + 1) * per]
        buf = torch.from_numpy(corpus.pack_sequences(mine, global_first=(rank == 0)).copy()) if mine else torch.empty(0, dtype=torch.uint8)
        symbols, d_rank = corpus.discover_alphabet(buf, dist.group.WORLD)
        out = {}
        for n in range(1, n_max + 1):
            g = data_builder.build_level_graph_partitioned(buf, n, symbols, d_rank, 1e-9, dist.group.WORLD)
            blk = {k: g.block[k].numpy() for k in ("rowptr", "col", "val_out", "val_in", "val_und", "in_src", "in_dst", "in_w")}
            blk["a_out"] = tuple(t.numpy() for t in g.a_out)
            out[n] = (g.node_sequences, g.number_of_nodes, g.number_of_edges, blk)
        q.put((rank, out))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,world", [("build_protein", 2), ("build_ragged", 3), ("build_ka1", 4)])
def test_fully_partitioned_build_equals_reference(name, world):
    """Count shards -> reduce-scatter over key ranges -> per-range extraction -> re-deal onto row blocks -> partitioned
    normalisation: the reassembled blocks are the reference's golden graph (nodes, A_out_w / A_in_w bit-exact, the three
    propagation matrices to 2e-7), at every n level, although no rank ever held the merged table or the whole graph."""
    import re
    from protgram_directgcn_b200.host.partitioned import row_range
    g = load(name)
    seqs = ["".join(c.split("\n")[1:]).upper() for c in str(g["fasta"]).split(">") if c.strip()]
    seqs = [s for s in (re.sub(r"\s+", "", s) for s in seqs) if s]
    n_max = BUILD_FIXTURES[name]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_partitioned_build_worker, args=(r, world, port, seqs, n_max, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert all(not isinstance(v, str) for v in res.values()), res
    for n in range(1, n_max + 1):
        nodes = list(g[f"n{n}_nodes"])
        N = len(nodes)
        cat = lambda key: np.concatenate([res[r][n][3][key] for r in range(world)])
        rows = []
        for r in range(world):
            names, num_nodes, num_edges, blk = res[r][n]
            assert names == nodes and num_nodes == N and num_edges == int(g[f"n{n}_number_of_edges"])
            lo, hi, per = row_range(N, r, world)
            rows.append(np.repeat(np.arange(lo, hi), np.diff(blk["rowptr"][: hi - lo + 1])))
            assert np.all((blk["a_out"][0] >= lo) & (blk["a_out"][0] < hi))
        rows = np.concatenate(rows)
        a_out = [np.concatenate([res[r][n][3]["a_out"][i] for r in range(world)]) for i in range(3)]
        order = np.lexsort((a_out[1], a_out[0]))
        assert np.array_equal(np.stack([a_out[0][order], a_out[1][order]]), g[f"n{n}_A_out_w_idx"])
        assert np.array_equal(a_out[2][order], g[f"n{n}_A_out_w_val"])
        assert np.array_equal(np.stack([cat("in_src"), cat("in_dst")]), g[f"n{n}_A_in_w_idx"])
        assert np.array_equal(cat("in_w"), g[f"n{n}_A_in_w_val"])
        for m, key in (("mathcal_A_out", "val_out"), ("mathcal_A_in", "val_in"), ("A_undirected_norm_sparse", "val_und")):
            assert np.array_equal(np.stack([rows, cat("col").astype(np.int64)]), g[f"n{n}_{m}_idx"]), (n, m)
            ref = g[f"n{n}_{m}_val"]
            assert np.max(np.abs(cat(key) - ref) / np.abs(ref)) <= 2e-7, (n, m)


def _model_worker(rank, world, port, payload, q):
    try:
        _init(rank, world, port)
        import protgram_directgcn_b200 as pg
        from protgram_directgcn_b200.host import partitioned as part
        src, dst, cnt, n, x, y, dims, classes, state, use_vec, tc = payload
        if tc:   # tensor-core branch of the layer (spec kernels): backward goes through the scaled fan-out with EXCHANGED gates
            from protgram_directgcn_b200.host import protgram_directgcn as model_mod
            from tests import kernel_spec
            model_mod.TC_MODE = "force"
            kernel_spec.pg_layer_gemm_fwd_tc_supported = lambda f_in, f_out: 1
        lo, hi, per = part.row_range(n, rank, world)
        mine = (src >= lo) & (src < hi)
        res = part.normalize_row_partitioned(torch.from_numpy(src[mine]), torch.from_numpy(dst[mine]),
                                             torch.from_numpy(cnt[mine].astype(np.float32)), n)
        torch.manual_seed(0)
        model = pg.ProtGramDirectGCN(dims, per, classes, 1, 0, 0, 0.0, use_vec)
        sd = {}
        for k, v in state.items():          # replicated parameters as they are, per-node parameters cut to this rank's rows
            if k.rsplit(".", 1)[-1] in part.PER_NODE_PARAMETERS:
                blk = torch.zeros((per,) + tuple(v.shape[1:]), dtype=v.dtype)
                blk[: hi - lo] = v[lo:hi]
                sd[k] = blk
            else:
                sd[k] = v
        model.load_state_dict(sd, strict=True)
        model.eval()
        data = part.partitioned_data(torch.from_numpy(x[lo:hi]), part.local_csr(res), n)
        logp, emb = model(data)
        yl = torch.full((per,), -100, dtype=torch.int64)
        yl[: hi - lo] = torch.from_numpy(y[lo:hi])
        loss = torch.nn.functional.nll_loss(logp, yl, reduction="sum") / n      # ignore_index=-100 masks the padding
        loss.backward()
        part.allreduce_replicated_grads(model)
        grads = {k: p.grad.numpy() for k, p in model.named_parameters() if p.grad is not None}
        q.put((rank, {"logp": logp.detach().numpy()[: hi - lo], "emb": emb.detach().numpy()[: hi - lo], "grads": grads}))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("use_vec,tc", [(True, False), (False, False), (True, True)])
def test_row_partitioned_model_equals_single_process(use_vec, tc, monkeypatch):
    """SURVEY 8(e) row 3 end to end: the unchanged ProtGramDirectGCN on each rank's row block (partitioned normalisation ->
    PartitionedStructure -> fused layers with all-gathered SpMM operands) against the same model on the whole graph in one
    process: log-probs and embeddings row for row, gradients of the replicated parameters after the all-reduce, gradients
    of the per-node parameters (constant, gate vectors) slice for slice.  Per-node gates exercise the gate exchange."""
    import protgram_directgcn_b200 as pg
    from oracle import graph_oracle
    from protgram_directgcn_b200 import _native as nat
    from protgram_directgcn_b200.host import partitioned as part
    from tests import kernel_spec
    n, world, dims, classes = 157, 2, [12, 16, 16, 8], 5
    src, dst, cnt = random_count_graph(n, seed=11, density=0.05)
    rng = np.random.default_rng(3)
    x = rng.standard_normal((n, dims[0])).astype(np.float32)
    y = rng.integers(0, classes, n)
    torch.manual_seed(1)
    full = pg.ProtGramDirectGCN(dims, n, classes, 1, 0, 0, 0.0, use_vec)
    with torch.no_grad():
        for k, p in full.named_parameters():    # non-trivial gates / biases
            if "C_" in k or "bias" in k:
                p.add_(0.3 * torch.randn_like(p))
    state = {k: v.clone() for k, v in full.state_dict().items()}
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    payload = (src, dst, cnt, n, x, y, dims, classes, state, use_vec, tc)
    procs = [ctx.Process(target=_model_worker, args=(r, world, port, payload, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert all(not isinstance(v, str) for v in res.values()), res
    # single-process truth on the whole graph (same spec kernels)
    kernel_spec.install(monkeypatch, nat)      # undone at the end of the test: later tests see the real library again
    mats = graph_oracle.normalise_all(src, dst, cnt, n)
    ei = torch.from_numpy(np.stack([mats["mathcal_A_in"][1], mats["mathcal_A_in"][0]]))     # (source = column, target = row)
    ew = [torch.from_numpy(mats[m][2]) for m in ("mathcal_A_in", "mathcal_A_out", "A_undirected_norm_sparse")]
    data = pg.Data(x=torch.from_numpy(x), edge_index_in=ei, edge_weight_in=ew[0], edge_index_out=ei, edge_weight_out=ew[1],
                   edge_index_undirected_norm=ei, edge_weight_undirected_norm=ew[2])
    full.eval()
    logp, emb = full(data)
    loss = torch.nn.functional.nll_loss(logp, torch.from_numpy(y), reduction="sum") / n
    loss.backward()
    got_logp = np.concatenate([res[r]["logp"] for r in range(world)])
    got_emb = np.concatenate([res[r]["emb"] for r in range(world)])
    assert np.max(np.abs(got_logp - logp.detach().numpy())) <= 2e-5
    assert np.max(np.abs(got_emb - emb.detach().numpy())) <= 2e-5
    for k, p in full.named_parameters():
        ref = p.grad.numpy() if p.grad is not None else None
        if ref is None:
            continue
        if k.rsplit(".", 1)[-1] in part.PER_NODE_PARAMETERS:
            got = np.concatenate([res[r]["grads"][k][: part.row_range(n, r, world)[1] - part.row_range(n, r, world)[0]] for r in range(world)])
        else:
            got = res[0]["grads"][k]
            assert np.array_equal(got, res[1]["grads"][k]), k          # all-reduced: identical on every rank
        assert np.max(np.abs(got - ref)) <= 2e-5 * max(1.0, float(np.max(np.abs(ref)))), k


def _struct_worker(rank, world, port, payload, q):
    try:
        _init(rank, world, port)
        from protgram_directgcn_b200.host import partitioned as part
        rowptr, col, vals, n, x, scales, dz, mode = payload
        if mode == "allgather":
            part.EXCHANGE_MODE = "allgather"       # round-1 exchange: every row to everybody
        elif mode is not None:
            part.PIPELINE_CHUNKS, part.PIPELINE_MIN_BYTES = int(mode), 0   # halo exchange pipelined over feature-column chunks
        lo, hi, per = part.row_range(n, rank, world)
        st = part.PartitionedStructure(part.slice_rows(rowptr, col, vals, lo, hi, per), n)
        pad = lambda t: torch.cat([t[lo:hi], torch.zeros((per - (hi - lo),) + tuple(t.shape[1:]), dtype=t.dtype)])
        z_vec = st.fanout(pad(x), x.shape[1], scales=tuple(pad(s) for s in scales), scale_stride=1)
        z_sca = st.fanout(pad(x), x.shape[1], scales=tuple(s[:1].clone() for s in scales), scale_stride=0)
        dx = st.fanin(pad(dz), x.shape[1], pad(x))
        q.put((rank, z_vec[: hi - lo].numpy(), z_sca[: hi - lo].numpy(), dx[: hi - lo].numpy()))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc(), None, None))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("limit", [None, 3, "allgather"])
def test_partitioned_structure_scaled_fanout_and_fanin(limit):
    """The two calls the tensor-core backward makes on a structure (fan-out with per-SOURCE-row gate scales, which must be
    exchanged; fan-in with a row-local init) against dense algebra, world size 3 with a short last block: halo exchange in
    one piece, pipelined over 3 feature-column chunks, and the round-1 all-gather exchange."""
    rng = np.random.default_rng(8)
    n, f, world = 100, 12, 3
    mask = rng.random((n, n)) < 0.06
    mask |= mask.T
    dense = [np.where(mask, rng.standard_normal((n, n)), 0).astype(np.float32) for _ in range(3)]
    dense = [(d + d.T) / 2 for d in dense]
    rows, cols = np.nonzero(mask)
    rowptr = torch.from_numpy(np.concatenate([[0], np.cumsum(np.bincount(rows, minlength=n))]).astype(np.int64))
    col = torch.from_numpy(cols.astype(np.int32))
    vals = [torch.from_numpy(d[rows, cols]) for d in dense]
    x = torch.from_numpy(rng.standard_normal((n, f)).astype(np.float32))
    dz = torch.from_numpy(rng.standard_normal((n, 3 * f)).astype(np.float32))
    scales = [torch.from_numpy(rng.standard_normal(n).astype(np.float32)) for _ in range(3)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_struct_worker, args=(r, world, port, (rowptr, col, vals, n, x, scales, dz, limit), q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=180) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    assert all(not isinstance(r[1], str) for r in res), res
    z_vec, z_sca, dx = (np.concatenate([r[i] for r in res]) for i in (1, 2, 3))
    xd = x.numpy().astype(np.float64)
    ref_vec = np.concatenate([d.astype(np.float64) @ (s.numpy().astype(np.float64)[:, None] * xd) for d, s in zip(dense, scales)], axis=1)
    ref_sca = np.concatenate([float(s[0]) * (d.astype(np.float64) @ xd) for d, s in zip(dense, scales)], axis=1)
    ref_dx = xd + sum(d.astype(np.float64) @ dz.numpy().astype(np.float64)[:, v * f:(v + 1) * f] for v, d in enumerate(dense))
    for got, ref in ((z_vec, ref_vec), (z_sca, ref_sca), (dx, ref_dx)):
        assert np.max(np.abs(got - ref)) <= 1e-5 * np.max(np.abs(ref))


def _bench_key_range_worker(rank, world, port, seqs, n, q):
    try:
        _init(rank, world, port)
        import types
        import bench
        from protgram_directgcn_b200.host import corpus, data_builder

        class _Ev:                                     # CUDA events do not exist here; the choreography is what is under test
            def __init__(self, enable_timing=True):
                pass

            def record(self):
                pass

            def elapsed_time(self, other):
                return 1.0

        torch.cuda.Event, torch.cuda.synchronize = _Ev, (lambda *a, **k: None)
        per = (len(seqs) + world - 1) // world
        mine = seqs[rank * per:(rank + 1) * per]
        buf = torch.from_numpy(corpus.pack_sequences(mine, global_first=(rank == 0)).copy())
        symbols, d_rank = corpus.discover_alphabet(buf, dist.group.WORLD)
        sigma = int(symbols.size)
        bins, short = data_builder.count_level(buf, n, d_rank, sigma)
        dist.all_reduce(bins)
        s32 = short.to(torch.int32)
        dist.all_reduce(s32, op=dist.ReduceOp.MAX)
        node_code, src, dst, cnt = data_builder.extract_level(bins, s32.to(torch.uint8), n, sigma)
        pipe = types.SimpleNamespace(db=data_builder, dev=torch.device("cpu"), rank=rank, world=world, group=dist.group.WORLD)
        out = bench._build_key_range_variant(pipe, dist, buf, n, d_rank, sigma, None, 2, int(node_code.numel()), int(src.numel()),
                                             int(cnt.sum()))
        q.put((rank, out))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_bench_key_range_build_variant_choreography():
    """bench.py's `key_range_variant` of the multi-GPU build legs at world size 3 (gloo, spec kernels, dummy CUDA events): every
    rank gets through the same collectives and the totals equal the replicated build's."""
    from oracle import ngram_oracle
    seqs = ngram_oracle.synth_sequences(0, 90, 40)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bench_key_range_worker, args=(r, 3, port, seqs, 2, q)) for r in range(3)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert all(isinstance(v, dict) for v in res.values()), res
    assert all(v["matches_replicated_build"] for v in res.values())
    assert sum(v["edges_on_this_rank"] for v in res.values()) > 0


def _corpus_to_model_worker(rank, world, port, seqs, n, dims, classes, state, x, q):
    try:
        _init(rank, world, port)
        import protgram_directgcn_b200 as pg
        from protgram_directgcn_b200.host import corpus, data_builder, partitioned as part
        per_seq = (len(seqs) + world - 1) // world
        mine = seqs[rank * per_seq:(rank + 1) * per_seq]
        buf = torch.from_numpy(corpus.pack_sequences(mine, global_first=(rank == 0)).copy())
        symbols, d_rank = corpus.discover_alphabet(buf, dist.group.WORLD)
        g = data_builder.build_level_graph_partitioned(buf, n, symbols, d_rank, 1e-9, dist.group.WORLD)
        lo, hi, per = g.lo, g.hi, g.per
        model = pg.ProtGramDirectGCN(dims, per, classes, n, 0, 0, 0.0, True)
        sd = {}
        for k, v in state.items():
            if k.rsplit(".", 1)[-1] in part.PER_NODE_PARAMETERS:
                blk = torch.zeros((per,) + tuple(v.shape[1:]), dtype=v.dtype)
                blk[: hi - lo] = v[lo:hi]
                sd[k] = blk
            else:
                sd[k] = v
        model.load_state_dict(sd, strict=True)
        model.eval()
        data = part.partitioned_data(torch.from_numpy(x[lo:hi]), part.local_csr(g.block), g.number_of_nodes)
        _, emb = model(data)
        q.put((rank, g.node_sequences, emb.detach().numpy()[: hi - lo]))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc(), None))
    finally:
        dist.destroy_process_group()


def test_corpus_to_embeddings_without_a_whole_graph_anywhere(monkeypatch):
    """The whole chain on 2 ranks -- corpus shards -> counts -> reduce-scatter over key ranges -> key-range extraction ->
    re-deal -> partitioned normalisation -> the unchanged model on row blocks -- against the single-process pipeline
    (GraphBuilder-style build_level_graph + the model on the whole graph): same nodes, same embeddings."""
    import protgram_directgcn_b200 as pg
    from oracle import ngram_oracle
    from protgram_directgcn_b200 import _native as nat
    from protgram_directgcn_b200.host import corpus, data_builder
    from tests import kernel_spec
    seqs = ngram_oracle.synth_sequences(0, 120, 45)
    n, dims, classes, world = 2, [10, 16, 8], 4, 2
    kernel_spec.install(monkeypatch, nat)
    buf = torch.from_numpy(corpus.pack_sequences(seqs).copy())
    symbols, d_rank = corpus.discover_alphabet(buf)
    whole = data_builder.build_level_graph(buf, n, symbols, d_rank, 1e-9)
    N = whole.number_of_nodes
    torch.manual_seed(2)
    full = pg.ProtGramDirectGCN(dims, N, classes, n, 0, 0, 0.0, True).eval()
    with torch.no_grad():
        for k, p in full.named_parameters():
            if "C_" in k or "bias" in k:
                p.add_(0.2 * torch.randn_like(p))
    x = np.random.default_rng(0).standard_normal((N, dims[0])).astype(np.float32)
    _, emb_ref = full(whole.gcn_data(torch.from_numpy(x), "cpu"))
    state = {k: v.clone() for k, v in full.state_dict().items()}
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_corpus_to_model_worker, args=(r, world, port, seqs, n, dims, classes, state, x, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=240) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    assert all(not isinstance(r[1], str) for r in res), res
    assert all(r[1] == whole.node_sequences for r in res)
    emb = np.concatenate([r[2] for r in res])
    assert emb.shape == tuple(emb_ref.shape) and np.max(np.abs(emb - emb_ref.detach().numpy())) <= 2e-5


def test_row_partitioned_normalisation_of_an_edgeless_graph_is_empty():
    """Nodes but no edges (every sequence exactly n residues long): the reference leaves all matrices empty
    (graph_utils.py:219-223), so do the blocks -- not the identity pattern the tagged-key sort would give."""
    z = np.zeros(0, dtype=np.int64)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_norm_worker, args=(r, 2, port, (z, z, z, 5, np.zeros((5, 4), dtype=np.float32)), q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert all(isinstance(v, dict) and v["pattern_nnz"] == 0 and not v["rowptr"].any() for v in res.values()), res
