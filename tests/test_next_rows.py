"""Rows f2 / f4 of SURVEY.md section 8 (protein pooling, feature hand-off, next_node labels): the host mirrors of the
reference functions against the reference-generated fixture -- on CPU with the native entry points swapped for their
executable specification, and on the GPU through the C ABI (bit-exact for the pooling)."""
import numpy as np
import pytest
import torch

import protgram_directgcn_b200 as pg
from protgram_directgcn_b200 import _native as nat
from oracle import next_oracle
from tests import kernel_spec
from tests.helpers import load


def _proteins(g):
    return list(zip([str(x) for x in g["prot_ids"]], [str(x) for x in g["prot_seqs"]]))


class _G:
    """Just what the trainer-side helpers read from a graph object."""

    def __init__(self, g, n):
        self.node_sequences = [str(x) for x in g[f"n{n}_nodes"]]
        self.number_of_nodes = len(self.node_sequences)
        self.node_to_idx = {s: i for i, s in enumerate(self.node_sequences)}
        self.A_out_w = torch.sparse_coo_tensor(torch.from_numpy(g[f"n{n}_a_out_idx"]), torch.from_numpy(g[f"n{n}_a_out_val"]),
                                               (self.number_of_nodes,) * 2).coalesce()


def check_all(g):
    seqs = _proteins(g)
    for n in (1, 2, 3):
        graph = _G(g, n)
        emb = g[f"n{n}_emb"]
        pooled = pg.EmbeddingProcessor.pool_ngram_embeddings_for_protein_fast(seqs, n, graph.node_to_idx, emb)
        assert list(pooled.keys()) == [str(x) for x in g[f"n{n}_pooled_ids"]]
        assert np.array_equal(np.stack(list(pooled.values())), g[f"n{n}_pooled"])        # bit exact vs the reference
        assert all(v.dtype == emb.dtype for v in pooled.values())
        labels, ncls = pg.generate_next_node_labels(graph)
        sets = next_oracle.next_node_label_sets(g[f"n{n}_a_out_idx"], g[f"n{n}_a_out_val"], graph.number_of_nodes)
        assert ncls == graph.number_of_nodes and labels.dtype == torch.int64 and not labels.is_cuda
        assert all(int(labels[i]) in set(sets[i].tolist()) for i in range(graph.number_of_nodes))
        assert all(int(labels[i]) == int(sets[i][0]) for i in range(graph.number_of_nodes))   # ties -> first maximal successor
        if n > 1:
            prev = _G(g, n - 1)
            x = pg.init_level_features(graph, prev.node_to_idx, g[f"n{n - 1}_emb"])
            assert np.array_equal(x.cpu().numpy(), g[f"n{n}_x_init_ref"])      # the reference's own loop, bit for bit
            assert np.array_equal(g[f"n{n}_x_init_ref"], g[f"n{n}_x_init_oracle"])
    # a shuffled previous-level id map must give the same features (rows follow the map, not the code order)
    prev = _G(g, 2)
    perm = np.random.default_rng(0).permutation(prev.number_of_nodes)
    shuffled = {s: int(perm[i]) for s, i in prev.node_to_idx.items()}
    emb_shuffled = np.empty_like(g["n2_emb"])
    emb_shuffled[perm] = g["n2_emb"]
    x = pg.init_level_features(_G(g, 3), shuffled, emb_shuffled)
    assert np.array_equal(x.cpu().numpy(), g["n3_x_init_ref"])
    assert pg.EmbeddingProcessor.pool_ngram_embeddings_for_protein_fast([], 2, {}, g["n2_emb"]) == {}
    with pytest.raises(ValueError):
        pg.EmbeddingProcessor.pool_ngram_embeddings_for_protein_fast([("p", "ACé")], 1, _G(g, 1).node_to_idx, g["n1_emb"])


def test_next_rows_host_logic_cpu(monkeypatch):
    kernel_spec.install(monkeypatch, nat)
    check_all(load("next_rows"))


@pytest.mark.gpu
def test_next_rows_gpu():
    check_all(load("next_rows"))


@pytest.mark.gpu
def test_pooling_long_and_repetitive_proteins_gpu():
    """Proteins longer than one bitmap chunk of distinct n-grams, homopolymers, n = 4 (194k-bit bitmap, 24 chunks)."""
    rng = np.random.default_rng(5)
    aa = list("ACDEFGHIKLMNPQRSTVWY")
    n = 4
    seqs = [("long", "".join(rng.choice(aa, size=20_000))), ("homo", "A" * 5000), ("short", "ACD"), ("one", "ACDE"),
            ("mix", "".join(rng.choice(aa[:3], size=3000)))]
    grams = sorted({s[i:i + n] for _, s in seqs for i in range(len(s) - n + 1)} | {"".join(x) for x in rng.choice(aa, size=(500, n))})
    ngram_map = {s: i for i, s in enumerate(grams)}
    emb = rng.standard_normal((len(grams), 96)).astype(np.float32)
    got = pg.EmbeddingProcessor.pool_ngram_embeddings_for_protein_fast(seqs, n, ngram_map, emb)
    ids, pooled, valid = next_oracle.pool_proteins(seqs, n, ngram_map, emb)
    assert list(got.keys()) == [p for p, v in zip(ids, valid) if v] and "short" not in got
    assert np.array_equal(np.stack(list(got.values())), pooled[valid])


class _PatternGraph:
    """Three value-symmetric matrices on one symmetric pattern with self loops (what a reference-built graph carries)."""

    def __init__(self, n, density, seed):
        rng = np.random.default_rng(seed)
        a = rng.random((n, n)) < density
        a = a | a.T | np.eye(n, dtype=bool)
        r, c = np.nonzero(a)
        idx = torch.from_numpy(np.stack([r, c]).astype(np.int64))
        self.number_of_nodes = n
        mats = []
        for _ in range(3):
            v = rng.standard_normal((n, n)).astype(np.float32)
            v = (v + v.T) / 2
            mats.append(torch.sparse_coo_tensor(idx, torch.from_numpy(v[r, c]), (n, n)).coalesce())
        self.mathcal_A_in, self.mathcal_A_out, self.A_undirected_norm_sparse = mats


def check_clusters(device):
    n = 300
    graph = _PatternGraph(n, 0.03, 1)
    rng = np.random.default_rng(2)
    perm = rng.permutation(n)
    clusters = [sorted(perm[:120].tolist()), sorted(perm[120:299].tolist()), [int(perm[299])], perm[:50].tolist(), []]
    full = pg.Data(x=torch.randn(n, 6), y=torch.arange(n))
    subs = pg.create_clustered_subgraphs(graph, clusters, full, device=device)
    assert len(subs) == len(clusters)
    for cl, d in zip(clusters, subs):
        assert torch.equal(d.original_indices.cpu(), torch.tensor(cl, dtype=torch.long))
        assert torch.equal(d.x.cpu(), full.x[cl]) and torch.equal(d.y.cpu(), full.y[cl])
        for name, m in (("in", graph.mathcal_A_in), ("out", graph.mathcal_A_out), ("undirected_norm", graph.A_undirected_norm_sparse)):
            ei_ref, w_ref = next_oracle.subgraph(cl, m.indices().numpy(), m.values().numpy(), n)
            ei = getattr(d, f"edge_index_{name}").cpu().numpy()
            w = getattr(d, f"edge_weight_{name}").cpu().numpy()
            assert ei.dtype == np.int64 and ei.shape == ei_ref.shape
            if cl == sorted(cl):                       # ascending cluster: the reference's edge order, bit for bit
                assert np.array_equal(ei, ei_ref) and np.array_equal(w, w_ref)
            else:                                      # same edge set, rows in cluster order
                key = lambda e: np.lexsort((e[1], e[0]))
                assert np.array_equal(ei[:, key(ei)], ei_ref[:, key(ei_ref)]) and np.array_equal(w[key(ei)], w_ref[key(ei_ref)])
    return graph, clusters, subs


def test_cluster_subgraphs_host_logic_cpu(monkeypatch):
    kernel_spec.install(monkeypatch, nat)
    check_clusters("cpu")


@pytest.mark.gpu
def test_cluster_subgraphs_gpu():
    graph, clusters, subs = check_clusters("cuda")
    # the layer on a cluster batch: the sub-CSR handed over by the extraction == the general path on cloned edge tensors
    from protgram_directgcn_b200.host import protgram_directgcn as model_mod
    torch.manual_seed(0)
    model = pg.ProtGramDirectGCN([6, 16, 8], graph.number_of_nodes, 4, 2, 0, 512, 0.0, True).cuda().eval()
    d = subs[0]
    with torch.no_grad():
        a, _ = model(data=d)
        model_mod._STRUCT_CACHE.clear()
        clone = pg.Data(**{k: (v.clone() if torch.is_tensor(v) else v) for k, v in vars(d).items()})
        b, _ = model(data=clone)
    assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
