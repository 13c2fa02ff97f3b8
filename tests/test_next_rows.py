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
            assert np.array_equal(x.cpu().numpy(), g[f"n{n}_x_init_oracle"])
    # a shuffled previous-level id map must give the same features (rows follow the map, not the code order)
    prev = _G(g, 2)
    perm = np.random.default_rng(0).permutation(prev.number_of_nodes)
    shuffled = {s: int(perm[i]) for s, i in prev.node_to_idx.items()}
    emb_shuffled = np.empty_like(g["n2_emb"])
    emb_shuffled[perm] = g["n2_emb"]
    x = pg.init_level_features(_G(g, 3), shuffled, emb_shuffled)
    assert np.array_equal(x.cpu().numpy(), g["n3_x_init_oracle"])
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
