"""Host logic above the C ABI, exercised on CPU by swapping the native entry points for their
executable specification (tests/kernel_spec.py).  What this covers: FASTA ingest, corpus packing,
alphabet ranking, node decoding, DirectedNgramGraph assembly, GraphBuilder.run() file contract,
DirectGCN parameter packing + autograd wiring, state_dict compatibility -- all against the
reference-generated goldens.  The CUDA kernels themselves are covered by the -m gpu tests."""
import os

import numpy as np
import pytest
import torch

import protgram_directgcn_b200 as pg
from protgram_directgcn_b200 import _native as nat
from protgram_directgcn_b200.host import protgram_directgcn as model_mod
from tests import kernel_spec
from tests.helpers import BUILD_FIXTURES, MATS, MODEL_FIXTURES, fasta_sequences, golden_edges, load, rel_err


@pytest.fixture
def spec_native(monkeypatch):
    kernel_spec.install(monkeypatch, nat)
    model_mod._STRUCT_CACHE.clear()
    yield
    model_mod._STRUCT_CACHE.clear()


def check_graph_against_golden(graph, g, n, val_tol=2e-7):
    assert graph.number_of_nodes == int(g[f"n{n}_number_of_nodes"])
    assert graph.number_of_edges == int(g[f"n{n}_number_of_edges"])
    assert graph.node_sequences == list(g[f"n{n}_nodes"])
    assert graph.node_to_idx == {s: i for i, s in enumerate(g[f"n{n}_nodes"])}
    assert graph.n_value == n
    for m in MATS:
        t = getattr(graph, m)
        assert t.is_sparse and t.is_coalesced() and t.dtype == torch.float32 and t.indices().dtype == torch.int64
        assert tuple(t.shape) == (graph.number_of_nodes,) * 2
        assert np.array_equal(t.indices().cpu().numpy(), g[f"n{n}_{m}_idx"]), (n, m)
        if m in ("A_out_w", "A_in_w"):
            assert np.array_equal(t.values().cpu().numpy(), g[f"n{n}_{m}_val"]), (n, m)  # counts: bit exact
        else:
            assert rel_err(t.values().cpu().numpy(), g[f"n{n}_{m}_val"]) <= val_tol, (n, m)


@pytest.mark.parametrize("name", sorted(BUILD_FIXTURES))
def test_graph_builder_run_matches_reference(name, tmp_path, spec_native):
    g = load(name)
    cfg = pg.Config()
    cfg.GCN_INPUT_FASTA_PATH = fasta_sequences(str(g["fasta"]), tmp_path)
    cfg.BASE_OUTPUT_DIR = tmp_path / "out"
    cfg.GRAPH_OBJECTS_DIR = cfg.BASE_OUTPUT_DIR / "1_graph_objects"
    cfg.GCN_NGRAM_MAX_N = BUILD_FIXTURES[name]
    cfg.GRAPH_BUILDER_WORKERS = 1
    pg.GraphBuilder(cfg).run()
    for n in range(1, BUILD_FIXTURES[name] + 1):
        path = cfg.GRAPH_OBJECTS_DIR / f"ngram_graph_n{n}.pkl"
        assert path.exists()  # the only thing the reference's own smoke test asserts (unit_tests.py:79-80)
        graph = pg.DataUtils.load_object(str(path))
        assert isinstance(graph, pg.DirectedNgramGraph)
        check_graph_against_golden(graph, g, n)


@pytest.mark.parametrize("chunk_bytes,resident", [(64, 1 << 40), (64, 0), (200, 300)])
def test_graph_builder_streamed_chunks_match_reference(chunk_bytes, resident, tmp_path, spec_native):
    """Corpora larger than HBM stream through in chunks cut at sequence boundaries (resident or
    re-uploaded per level); the tables accumulate and the graphs are those of the one-buffer build."""
    g = load("build_ragged")
    cfg = pg.Config()
    cfg.GCN_INPUT_FASTA_PATH = fasta_sequences(str(g["fasta"]), tmp_path)
    cfg.BASE_OUTPUT_DIR = tmp_path / "out"
    cfg.GRAPH_OBJECTS_DIR = cfg.BASE_OUTPUT_DIR / "1_graph_objects"
    cfg.GCN_NGRAM_MAX_N = BUILD_FIXTURES["build_ragged"]
    cfg.GRAPH_BUILDER_CHUNK_BYTES = chunk_bytes
    cfg.GRAPH_BUILDER_RESIDENT_BYTES = resident
    pg.GraphBuilder(cfg).run()
    for n in range(1, BUILD_FIXTURES["build_ragged"] + 1):
        check_graph_against_golden(pg.DataUtils.load_object(str(cfg.GRAPH_OBJECTS_DIR / f"ngram_graph_n{n}.pkl")), g, n)


def test_stream_chunks_partition_the_corpus():
    from protgram_directgcn_b200.host import corpus
    seqs = ["ACDE", "", "K", "LMNPQ" * 7, "RST", "VW", "Y" * 40] * 5
    whole = corpus.pack_sequences(seqs).tobytes()
    for world in (1, 2, 3):
        pieces = []
        for rank in range(world):
            chunks = list(corpus.stream_chunks(iter(seqs), chunk_bytes=50, rank=rank, world=world, block=2))
            assert all(c.size and c[-1] == corpus.SEP for c in chunks)          # cut at sequence boundaries only
            pieces += [p for c in chunks for p in c.tobytes().split(b"\xff")[:-1]]
        assert sorted(pieces) == sorted(whole.split(b"\xff")[:-1])              # same multiset of padded sequences
        assert sum(p.startswith(b" ") and len(p) > 1 for p in pieces) == 1                      # one leading space: global sequence #0
    with pytest.raises(ValueError):
        list(corpus.stream_chunks(iter(["AC\u00e9"])))


def test_directed_ngram_graph_from_parquet_unsorted(tmp_path, spec_native):
    """Constructor contract of reference graph_utils.py:91-125: nodes dict + parquet edge table in
    arbitrary row order."""
    import pandas as pd
    g = load("build_protein")
    n = 2
    idx = g[f"n{n}_A_out_w_idx"]
    perm = np.random.default_rng(0).permutation(idx.shape[1])
    df = pd.DataFrame({"source": idx[0][perm], "target": idx[1][perm],
                       "weight": g[f"n{n}_A_out_w_val"][perm].astype(np.int64)})
    path = str(tmp_path / "edges.parquet")
    df.to_parquet(path, index=False)
    nodes = dict(enumerate(g[f"n{n}_nodes"].tolist()))
    graph = pg.DirectedNgramGraph(nodes, path, epsilon_propagation=1e-9, n_value=n)
    check_graph_against_golden(graph, g, n)
    # the trainer re-creates the propagation matrices after loading (protgram_directgcn_trainer.py:299)
    graph.mathcal_A_out = None
    graph._create_propagation_matrices_for_gcn()
    check_graph_against_golden(graph, g, n)


def test_empty_and_missing_inputs(tmp_path, spec_native, capsys):
    g = pg.DirectedNgramGraph({0: "A", 1: "C"}, None)
    assert g.number_of_nodes == 2 and g.number_of_edges == 0
    for m in MATS:
        assert getattr(g, m)._nnz() == 0 and tuple(getattr(g, m).shape) == (2, 2)
    g0 = pg.DirectedNgramGraph({}, None)
    assert g0.number_of_nodes == 0
    cfg = pg.Config()
    cfg.GCN_INPUT_FASTA_PATH = tmp_path / "nope.fasta"
    cfg.BASE_OUTPUT_DIR = tmp_path / "o"
    cfg.GRAPH_OBJECTS_DIR = tmp_path / "o" / "g"
    pg.GraphBuilder(cfg).run()  # prints, never raises (reference data_builder.py:110-112)
    assert "not found" in capsys.readouterr().out
    empty = tmp_path / "empty.fasta"
    empty.write_text(">only_header\n\n")
    cfg.GCN_INPUT_FASTA_PATH = empty
    pg.GraphBuilder(cfg).run()
    assert "No sequences found" in capsys.readouterr().out


def _load_model(g):
    dims = [int(d) for d in g["dims"]]
    model = pg.ProtGramDirectGCN(layer_dims=dims, num_graph_nodes=int(g["num_graph_nodes"]),
                                 task_num_output_classes=int(g["num_classes"]), n_gram_len=int(g["n_gram_len"]),
                                 one_gram_dim=int(g["one_gram_dim"]), max_pe_len=16, dropout=0.0,
                                 use_vector_coeffs=bool(g["use_vec"]))
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd:")}
    assert set(sd) == set(model.state_dict()), "state_dict keys must equal the reference's"
    model.load_state_dict(sd, strict=True)
    return model


def _data(g, device="cpu"):
    kw = {"x": torch.from_numpy(g["x"]).clone().to(device).requires_grad_(True)}
    for k, name in (("in", "in"), ("out", "out"), ("und", "undirected_norm")):
        kw[f"edge_index_{name}"] = torch.from_numpy(g[f"ei_{k}"]).to(device)
        kw[f"edge_weight_{name}"] = torch.from_numpy(g[f"ew_{k}"]).to(device) if f"ew_{k}" in g.files else None
    if "original_indices" in g.files:
        kw["original_indices"] = torch.from_numpy(g["original_indices"]).to(device)
    return pg.Data(**kw)


def run_model_case(g, device, tol_fwd, tol_bwd):
    model = _load_model(g).to(device)
    data = _data(g, device)
    model.eval()
    # per-layer embeddings: the golden holds the reference's conv outputs; the stack applies
    # leaky_relu(conv + res_proj(h)) on top (protgram_directgcn.py:210-215)
    with torch.no_grad():
        _, layers = model.embed(data, return_layers=True)
        sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd:")}
        h_ref = torch.from_numpy(g["x"])
        if "pe_layer.weight" in sd and int(g["one_gram_dim"]) > 0 and h_ref.shape[1] == int(g["n_gram_len"]) * int(g["one_gram_dim"]):
            k = min(int(g["n_gram_len"]), sd["pe_layer.weight"].shape[0])
            h_ref = h_ref.clone().view(-1, int(g["n_gram_len"]), int(g["one_gram_dim"]))
            h_ref[:, :k, :] += sd["pe_layer.weight"][:k].unsqueeze(0)
            h_ref = h_ref.view(g["x"].shape[0], -1)
        for i, h_mine in enumerate(layers):
            res = h_ref @ sd[f"res_projs.{i}.weight"].t() + sd[f"res_projs.{i}.bias"] if f"res_projs.{i}.weight" in sd else h_ref
            h_ref = torch.nn.functional.leaky_relu(torch.from_numpy(g[f"layer{i}_out"]) + res)
            assert rel_err(h_mine.cpu().numpy(), h_ref.numpy()) <= tol_fwd, f"layer {i}"
    logp, emb = model(data=data)
    assert rel_err(logp.detach().cpu().numpy(), g["logp"]) <= tol_fwd
    assert rel_err(emb.detach().cpu().numpy(), g["emb"]) <= tol_fwd
    y = torch.from_numpy(g["y"]).to(device)
    wvec = torch.from_numpy(g["wvec"]).to(device)
    loss = torch.nn.functional.nll_loss(logp, y) + (emb * wvec).sum()
    assert abs(float(loss) - float(g["loss"])) <= tol_fwd * max(1.0, abs(float(g["loss"])))
    loss.backward()
    assert rel_err(data.x.grad.cpu().numpy(), g["grad_x"]) <= tol_bwd
    # row f1: the fused decoder-output + log_softmax + nll path must give the reference's loss and gradient too
    grads_unfused = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    model.zero_grad(set_to_none=True)
    data2 = _data(g, device)
    h = model.embed(data2)
    emb2 = pg.EmbeddingProcessor.l2_normalize_torch(h, eps=model.l2_eps)
    last = model.decoder_fc[-1]
    loss2 = model_mod.linear_log_softmax_nll(model.decoder_fc[:-1](h), last.weight, last.bias, y) + (emb2 * wvec).sum()
    assert abs(float(loss2) - float(g["loss"])) <= tol_fwd * max(1.0, abs(float(g["loss"])))
    loss2.backward()
    assert rel_err(data2.x.grad.cpu().numpy(), g["grad_x"]) <= tol_bwd
    for k, p in model.named_parameters():
        if k in grads_unfused:
            assert rel_err(p.grad.cpu().numpy(), grads_unfused[k].cpu().numpy()) <= tol_bwd, k
    for k, p in model.named_parameters():
        ref = g["grad:" + k]
        got = p.grad.cpu().numpy() if p.grad is not None else np.zeros_like(ref)
        assert rel_err(got, ref) <= tol_bwd or float(np.max(np.abs(ref))) < 1e-12, k
    return model


@pytest.mark.parametrize("name", MODEL_FIXTURES)
def test_model_host_logic_matches_reference(name, spec_native):
    run_model_case(load(name), "cpu", 1e-5, 1e-4)


def test_directgcn_layer_alone_matches_oracle(spec_native):
    """DirectGCNLayer.forward on its own (no residual / activation), reference :93-135."""
    from oracle import directgcn_oracle
    g = load("model_general_scalar")
    model = _load_model(g)
    data = _data(g)
    layer = model.convs[0]
    out = layer(data.x, data.edge_index_in, data.edge_weight_in, data.edge_index_out, data.edge_weight_out,
                data.edge_index_undirected_norm, data.edge_weight_undirected_norm)
    p = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd:")}
    ref = directgcn_oracle.directgcn_layer(p, "convs.0.", data.x.detach(), data.edge_index_in, data.edge_weight_in,
                                           data.edge_index_out, data.edge_weight_out, data.edge_index_undirected_norm,
                                           data.edge_weight_undirected_norm)
    assert rel_err(out.detach().numpy(), ref.numpy()) <= 1e-5


def test_fused_loss_ignores_out_of_range_labels(spec_native):
    """ignore_index semantics of the fused loss (row f1) == F.nll_loss(..., ignore_index=-100)."""
    g = load("model_refgraph")
    model = _load_model(g)
    model.eval()
    y = torch.from_numpy(g["y"]).clone()
    y[::3] = -100
    ref = torch.nn.functional.nll_loss(model(data=_data(g))[0], y)
    ref.backward()
    grads_ref = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    model.zero_grad(set_to_none=True)
    out = model.nll_loss(_data(g), y, has_ignored=True)
    out.backward()
    assert abs(float(out) - float(ref)) <= 1e-6 * max(1.0, abs(float(ref)))
    for k, p in model.named_parameters():
        if k in grads_ref:
            assert rel_err(p.grad.numpy(), grads_ref[k].numpy()) <= 1e-5, k


def test_missing_inputs_raise_value_error(spec_native):
    g = load("model_cluster_batch")
    model = _load_model(g)
    with pytest.raises(ValueError):
        model(data=pg.Data(x=torch.zeros(3, 10)))
    with pytest.raises(ValueError):
        pg.ProtGramDirectGCN([4], 3, 2, 1, 0, 0, 0.0, True)


def test_no_cpu_fallback():
    """Without the test-only spec, CPU tensors must raise: the product has no CPU path."""
    g = load("model_cluster_batch")
    model = _load_model(g)
    with pytest.raises(nat.NativeError):
        model(data=_data(g))
    if not torch.cuda.is_available():
        with pytest.raises(nat.NativeError):
            pg.DirectedNgramGraph.from_edge_arrays({0: "A", 1: "C"}, np.array([0]), np.array([1]), np.array([2.0]))


def test_graph_pickle_csr_sidecar_roundtrip(tmp_path, spec_native):
    """Row f3, on-disk hand-off: save_object writes the reference-compatible pickle plus `<path>.csr.npz`; load_object
    attaches the sidecar when it matches, gcn_data() then hands the layer the CSR without sorting edge lists; a stale or
    foreign sidecar is ignored; the pickle alone still loads."""
    import pickle
    import shutil
    g = load("build_protein")
    src, dst, w = golden_edges(g, 2)
    nodes = {i: s for i, s in enumerate(g["n2_nodes"])}
    graph = pg.DirectedNgramGraph.from_edge_arrays(nodes, src, dst, w.astype(np.float32), n_value=2, assume_coalesced=True)
    path = str(tmp_path / "g.pkl")
    pg.DataUtils.save_object(graph, path)
    side = path + pg.DataUtils.SIDECAR_SUFFIX
    assert os.path.exists(side)
    z = np.load(side)
    p = graph.mathcal_A_in._nnz()
    assert z["rowptr"].dtype == np.int64 and z["col"].dtype == np.int32 and z["col"].shape == (p,) and int(z["rowptr"][-1]) == p
    with open(path, "rb") as fh:                       # the pickle knows nothing about the sidecar
        raw = pickle.load(fh)
    assert "_pg_sidecar" not in raw.__dict__ and "_pg_device" not in raw.__dict__
    loaded = pg.DataUtils.load_object(path)
    assert loaded.__dict__.get("_pg_sidecar") is not None
    check_graph_against_golden(loaded, g, 2)
    csr = loaded.propagation_csr_host()
    for m, k in (("mathcal_A_in", "val_in"), ("mathcal_A_out", "val_out"), ("A_undirected_norm_sparse", "val_und")):
        assert np.array_equal(csr[k], getattr(loaded, m).values().numpy())
    assert np.array_equal(csr["col"], loaded.mathcal_A_in.indices()[1].numpy().astype(np.int32))
    # the layer gets the CSR straight from the sidecar: no edge-list sort (pg_edges_to_csr is never called), same outputs
    torch.manual_seed(0)
    n_nodes = loaded.number_of_nodes
    model = pg.ProtGramDirectGCN([6, 8, 4], n_nodes, 3, 2, 0, 0, 0.0, True).eval()
    x = torch.randn(n_nodes, 6)
    calls = []
    real_call = nat.call
    nat.call = lambda name, *a: (calls.append(name), real_call(name, *a))[1]
    try:
        out_side = model(loaded.gcn_data(x, "cpu"))[1]
        assert "pg_edges_to_csr" not in calls and "pg_coo_from_csr" in calls
        del calls[:]
        raw_data = raw.gcn_data(x, "cpu")                       # no sidecar, no device arrays: the COO route
        out_coo = model(raw_data)[1]
        assert "pg_edges_to_csr" in calls
    finally:
        nat.call = real_call
    out_built = model(graph.gcn_data(x, "cpu"))[1]
    assert torch.equal(out_side, out_built) and rel_err(out_coo.detach().numpy(), out_built.detach().numpy()) <= 1e-6
    # a graph WITHOUT the device arrays (as after unpickling elsewhere) derives the same sidecar from its COO tensors
    raw.__dict__.pop("_pg_sidecar", None)
    derived = raw.propagation_csr_host()
    assert all(np.array_equal(derived[k], csr[k]) for k in ("rowptr", "col", "val_in", "val_out", "val_und"))
    # re-saving a loaded graph keeps the sidecar; a foreign sidecar is refused with a warning, not an error
    pg.DataUtils.save_object(loaded, str(tmp_path / "again.pkl"))
    assert os.path.exists(str(tmp_path / "again.pkl") + pg.DataUtils.SIDECAR_SUFFIX)
    src1, dst1, w1 = golden_edges(g, 1)
    other = pg.DirectedNgramGraph.from_edge_arrays({i: s for i, s in enumerate(g["n1_nodes"])}, src1, dst1, w1.astype(np.float32),
                                                   n_value=1, assume_coalesced=True)
    path1 = str(tmp_path / "g1.pkl")
    pg.DataUtils.save_object(other, path1)
    shutil.copy(side, path1 + pg.DataUtils.SIDECAR_SUFFIX)
    os.utime(path1 + pg.DataUtils.SIDECAR_SUFFIX, None)
    assert pg.DataUtils.load_object(path1).__dict__.get("_pg_sidecar") is None
    os.remove(side)
    assert pg.DataUtils.load_object(path).__dict__.get("_pg_sidecar") is None


@pytest.mark.parametrize("fixture", ["model_refgraph", "model_general_scalar"])
def test_tensor_core_path_host_wiring_matches_simt_path(spec_native, monkeypatch, fixture):
    """The tensor-core branch of the fused layer (forward TC GEMM, weight gradient, gate gradients from the dot-product epilogue,
    input gradient through the SCALED FAN-OUT + one GEMM instead of the fan-in) wired on the executable spec: same outputs and
    gradients as the SIMT branch, with and without residual projections, vector and scalar gates.  model_general_scalar carries
    three different UNSYMMETRIC edge lists (benchmarker contract): the regrouped input gradient must run over the
    source-grouped CSRs there."""
    for use_vec in (True, False):
        g = load(fixture)
        outs = {}
        for mode, bwd in (("off", "fanin"), ("off", "fanout"), ("force", "fanout")):
            monkeypatch.setattr(model_mod, "TC_MODE", mode)
            monkeypatch.setattr(model_mod, "BWD_DX_MODE", bwd)
            monkeypatch.setattr(kernel_spec, "pg_layer_gemm_fwd_tc_supported", lambda f_in, f_out: 1)
            torch.manual_seed(0)
            n = int(g["num_graph_nodes"])
            model = pg.ProtGramDirectGCN([8, 16, 16, 4], n, 3, 1, 0, 0, 0.0, use_vec).eval()
            with torch.no_grad():
                for k, p in model.named_parameters():
                    if "C_" in k or "bias" in k:
                        p.add_(0.3 * torch.randn_like(p))
            data = _data(g)
            data.x = torch.randn(n, 8, generator=torch.Generator().manual_seed(1)).requires_grad_(True)
            calls = []
            real_call = nat.call
            monkeypatch.setattr(nat, "call", lambda name, *a, _c=calls, _r=real_call: (_c.append(name), _r(name, *a))[1])
            logp, emb = model(data)
            (logp.sum() + (emb * emb).sum()).backward()
            monkeypatch.setattr(nat, "call", real_call)
            outs[(mode, bwd)] = (logp.detach(), emb.detach(), data.x.grad.clone(), {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}, calls)
        legacy, simt, tc = outs[("off", "fanin")], outs[("off", "fanout")], outs[("force", "fanout")]
        assert "pg_spmm_fanout_scaled" in tc[4] and "pg_layer_gate_grad_tc" in tc[4] and "pg_spmm_fanin" not in tc[4]
        assert "pg_spmm_fanout_scaled" in simt[4] and "pg_layer_gate_grad" in simt[4] and "pg_layer_gemm_bwd_dx" in simt[4]
        assert "pg_spmm_fanin" not in simt[4] and "pg_layer_gemm_fwd_tc" not in simt[4]
        assert "pg_spmm_fanin" in legacy[4] and "pg_layer_gemm_bwd_data" in legacy[4]
        for other in (simt, tc):
            for a, b in zip(legacy[:3], other[:3]):
                assert rel_err(b.numpy(), a.numpy()) <= 2e-5
            for k, v in legacy[3].items():
                assert rel_err(other[3][k].numpy(), v.numpy()) <= 5e-5, k


def test_preregistered_structure_survives_cache_pressure(spec_native):
    """a structure registered for a placeholder edge_index (CUDA-graph step, cluster sub-CSRs) rides on the
    tensor itself: however many other graphs pass through the LRU cache in between, the layer finds it again; and a
    placeholder WITHOUT its structure raises instead of silently training on one self-loop."""
    g = load("model_refgraph")
    n = int(g["num_graph_nodes"])
    ei_real = torch.from_numpy(g["ei_in"]) if "ei_in" in g.files else _data(g).edge_index_in
    d = _data(g)
    st_real = model_mod.get_structure((d.edge_index_in, d.edge_index_out, d.edge_index_undirected_norm),
                                      (d.edge_weight_in, d.edge_weight_out, d.edge_weight_undirected_norm), n)
    csr = st_real.by_dst[0]
    ph = torch.zeros((2, 1), dtype=torch.int64)
    ph._pg_placeholder = True
    vals = tuple(v.clone() for v in csr.vals)
    model_mod.register_symmetric_structure(ph, vals, n, csr.rowptr, csr.col)
    for i in range(model_mod.STRUCT_CACHE_ENTRIES + 8):       # cache pressure: > capacity other structures
        e = torch.tensor([[0, i % n], [i % n, 0]], dtype=torch.int64)
        model_mod.get_structure((e, e, e), (None, None, None), n)
    assert len(model_mod._STRUCT_CACHE) <= model_mod.STRUCT_CACHE_ENTRIES
    got = model_mod.get_structure((ph, ph, ph), vals, n)
    assert got.by_dst[0].col is csr.col and got.shared
    naked = torch.zeros((2, 1), dtype=torch.int64)
    naked._pg_placeholder = True
    with pytest.raises(RuntimeError):
        model_mod.get_structure((naked, naked, naked), vals, n)
    # same tensor, other weights: the tag does not apply (a fresh structure is built from the edge list instead)
    other = tuple(v.clone() for v in vals)
    assert model_mod.get_structure((ei_real, ei_real, ei_real), (None, None, None), n) is not got
