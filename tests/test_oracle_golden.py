"""The oracle is pinned here: every oracle function against the reference-generated goldens
(tests/golden/make_golden.py ran the reference's own modules).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import c_oracle, directgcn_oracle, graph_oracle, ngram_oracle
from tests.helpers import BUILD_FIXTURES, MATS, MODEL_FIXTURES, fasta_sequences, load, rel_err


def _seqs(g, tmp_path):
    return [s for _, s in ngram_oracle.parse_fasta(fasta_sequences(str(g["fasta"]), tmp_path))]


@pytest.mark.parametrize("name", sorted(BUILD_FIXTURES))
@pytest.mark.parametrize("impl", ["py", "np", "c"])
def test_builder_oracle_bit_exact(name, impl, tmp_path):
    g = load(name)
    seqs = _seqs(g, tmp_path)
    n_max = BUILD_FIXTURES[name]
    if impl == "c":
        buf = c_oracle.pack_corpus(seqs)
        symbols, rank = c_oracle.alphabet(buf)
    else:
        levels = ngram_oracle.build_all_levels(seqs, n_max, impl=impl)
    for n in range(1, n_max + 1):
        if impl == "c":
            bins, present = c_oracle.count_level(buf, n, rank, symbols.size)
            nodes, src, dst, w = c_oracle.bins_to_graph(bins, present, symbols, n)
            assert int(bins.sum()) == int(g[f"n{n}_n_transitions"])
        else:
            nodes, src, dst, w = levels[n]
        assert nodes == list(g[f"n{n}_nodes"])
        idx = g[f"n{n}_A_out_w_idx"]
        assert np.array_equal(src, idx[0]) and np.array_equal(dst, idx[1])
        assert np.array_equal(w.astype(np.float32), g[f"n{n}_A_out_w_val"])
        assert int(w.sum()) == int(g[f"n{n}_n_transitions"])


def test_known_answers_ka1_ka2():
    """SURVEY.md 8(c) hand-derived answers for the reference's two demo FASTAs."""
    g = load("build_ka1")
    assert list(g["n1_nodes"]) == [" ", "A", "C", "G", "T"]
    dense = np.zeros((5, 5))
    dense[g["n1_A_out_w_idx"][0], g["n1_A_out_w_idx"][1]] = g["n1_A_out_w_val"]
    assert dense.tolist() == [[0, 1, 0, 0, 0], [1, 0, 3, 2, 1], [0, 0, 0, 2, 1], [0, 2, 0, 0, 2], [2, 3, 0, 0, 2]]
    assert int(g["n1_n_transitions"]) == 22 and int(g["n2_n_transitions"]) == 19 and int(g["n3_n_transitions"]) == 16
    assert list(g["n2_nodes"]) == [" A", "A ", "AC", "AG", "AT", "CG", "CT", "GA", "GT", "T ", "TA", "TT"]
    assert list(g["n3_nodes"][:6]) == [" AC", "ACG", "ACT", "AGA", "ATA", "CGT"]
    g2 = load("build_ka2")
    got = {(int(s), int(d)): int(w) for s, d, w in zip(*g2["n1_A_out_w_idx"], g2["n1_A_out_w_val"])}
    S, A, C, G, T = range(5)
    assert got == {(S, A): 1, (A, S): 1, (A, C): 2, (A, G): 1, (C, S): 1, (C, G): 1, (G, A): 1, (G, T): 1,
                   (T, S): 1, (T, A): 1, (T, T): 1}


@pytest.mark.parametrize("name", sorted(BUILD_FIXTURES))
def test_graph_oracle_vs_reference(name):
    g = load(name)
    for n in range(1, BUILD_FIXTURES[name] + 1):
        N = int(g[f"n{n}_number_of_nodes"])
        idx = g[f"n{n}_A_out_w_idx"]
        # feed the edge table in a scrambled order: the result must not depend on it
        perm = np.random.default_rng(n).permutation(idx.shape[1])
        mats = graph_oracle.normalise_all(idx[0][perm], idx[1][perm],
                                          g[f"n{n}_A_out_w_val"][perm].astype(np.int64), N)
        for m in MATS:
            r, c, v = mats[m]
            assert np.array_equal(np.stack([r, c]), g[f"n{n}_{m}_idx"]), (name, n, m)
            # integer counts < 2^24: every fp32 step is reproducible -> expect (near) bit equality
            assert rel_err(v, g[f"n{n}_{m}_val"]) <= 2e-7, (name, n, m)
        # shared symmetric pattern
        assert np.array_equal(g[f"n{n}_mathcal_A_out_idx"], g[f"n{n}_mathcal_A_in_idx"])
        assert np.array_equal(g[f"n{n}_mathcal_A_out_idx"], g[f"n{n}_A_undirected_norm_sparse_idx"])


def _params(g, requires_grad=False):
    p = {k[3:]: torch.from_numpy(g[k]).clone() for k in g.files if k.startswith("sd:")}
    if requires_grad:
        for v in p.values():
            v.requires_grad_(True)
    return p


def _edges(g):
    out = []
    for k in ("in", "out", "und"):
        out.append(torch.from_numpy(g[f"ei_{k}"]))
        out.append(torch.from_numpy(g[f"ew_{k}"]) if f"ew_{k}" in g.files else None)
    return out


@pytest.mark.parametrize("name", MODEL_FIXTURES)
def test_directgcn_oracle_vs_reference(name):
    g = load(name)
    p = _params(g, requires_grad=True)
    x = torch.from_numpy(g["x"]).clone().requires_grad_(True)
    oi = torch.from_numpy(g["original_indices"]) if "original_indices" in g.files else None
    logp, emb, layers = directgcn_oracle.protgram_forward(
        p, x, *_edges(g), n_gram_len=int(g["n_gram_len"]), one_gram_dim=int(g["one_gram_dim"]),
        original_indices=oi, return_layers=True)
    for i, lo in enumerate(layers):
        assert rel_err(lo.detach().numpy(), g[f"layer{i}_out"]) <= 1e-6
    assert rel_err(logp.detach().numpy(), g["logp"]) <= 1e-6
    assert rel_err(emb.detach().numpy(), g["emb"]) <= 1e-6
    loss = torch.nn.functional.nll_loss(logp, torch.from_numpy(g["y"])) + (emb * torch.from_numpy(g["wvec"])).sum()
    loss.backward()
    assert rel_err(x.grad.numpy(), g["grad_x"]) <= 1e-5
    for k, v in p.items():
        ref = g["grad:" + k]
        got = v.grad.numpy() if v.grad is not None else np.zeros_like(ref)
        assert rel_err(got, ref) <= 1e-5 or np.max(np.abs(ref)) < 1e-12, k


def test_synth_corpus_is_shard_independent():
    a = ngram_oracle.synth_residues(0, 64, 50)
    b = np.concatenate([ngram_oracle.synth_residues(0, 40, 50), ngram_oracle.synth_residues(40, 24, 50)])
    assert np.array_equal(a, b)
    assert set(bytes(a.ravel()).decode()) <= set(ngram_oracle.AA)
