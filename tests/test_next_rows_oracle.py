"""Oracle for the 'next' rows of SURVEY.md section 8 (f2 pooling / feature hand-off, f4 labels) against the fixture
produced by the reference's own functions (tests/golden/make_golden_next.py)."""
import numpy as np

from oracle import next_oracle
from tests.helpers import load


def _proteins(g):
    return list(zip([str(x) for x in g["prot_ids"]], [str(x) for x in g["prot_seqs"]]))


def test_pooling_oracle_is_bit_exact_vs_reference():
    g = load("next_rows")
    seqs = _proteins(g)
    for n in (1, 2, 3):
        nodes = [str(x) for x in g[f"n{n}_nodes"]]
        ids, pooled, valid = next_oracle.pool_proteins(seqs, n, {s: i for i, s in enumerate(nodes)}, g[f"n{n}_emb"])
        kept = [p for p, v in zip(ids, valid) if v]
        assert kept == [str(x) for x in g[f"n{n}_pooled_ids"]]          # proteins without a known n-gram are dropped
        assert np.array_equal(pooled[valid], g[f"n{n}_pooled"])           # distinct n-grams, ascending id, fp32: bit exact
    assert "EMPTY" not in kept and "ONLYX" not in kept


def test_label_oracle_admits_the_reference_labels():
    g = load("next_rows")
    for n in (1, 2, 3):
        num = g[f"n{n}_nodes"].shape[0]
        sets = next_oracle.next_node_label_sets(g[f"n{n}_a_out_idx"], g[f"n{n}_a_out_val"], num)
        ref = g[f"n{n}_labels_ref"]
        assert len(sets) == num == ref.shape[0]
        assert all(int(ref[i]) in set(sets[i].tolist()) for i in range(num))
        single = [i for i in range(num) if sets[i].size == 1]
        assert all(int(ref[i]) == int(sets[i][0]) for i in single) and len(single) > 0


def test_feature_handoff_oracle_is_bit_exact_vs_reference_loop():
    """`n{n}_x_init_ref` was produced by executing the reference's own loop (protgram_directgcn_trainer.py:323-330, cut out of
    the module's source by make_golden_next.py)."""
    g = load("next_rows")
    for n in (2, 3):
        nodes = [str(x) for x in g[f"n{n}_nodes"]]
        prev = {str(s): i for i, s in enumerate(g[f"n{n - 1}_nodes"])}
        x = next_oracle.init_level_features(nodes, prev, g[f"n{n - 1}_emb"])
        assert np.array_equal(x, g[f"n{n}_x_init_ref"])


def test_subgraph_oracle_matches_the_published_pyg_example():
    """`torch_geometric.utils.subgraph` cannot run here (PyG is an unpinned, absent dependency of the reference), so the
    restatement is held against the one known answer PyG itself publishes: the example of the function's docstring
    (torch_geometric/utils/subgraph.py, `subgraph`): a path 0-1-...-6 stored in both directions, subset {3, 4, 5}
    -> edges (3,4) (4,3) (4,5) (5,4) with attributes 7, 8, 9, 10; with relabel_nodes=True (what the reference's call site
    protgram_directgcn_trainer.py:183-186 passes) the same edges renumbered 0..2.  A partial pin: one vector, not a fixture set."""
    edge_index = np.array([[0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6],
                           [1, 0, 2, 1, 3, 2, 4, 3, 5, 4, 6, 5]])
    edge_attr = np.arange(1, 13, dtype=np.float32)
    ei, ea = next_oracle.subgraph([3, 4, 5], edge_index, edge_attr, 7)
    assert np.array_equal(ea, np.array([7, 8, 9, 10], dtype=np.float32))
    assert np.array_equal(ei, np.array([[0, 1, 1, 2], [1, 0, 2, 1]]))          # relabelled: 3 -> 0, 4 -> 1, 5 -> 2
    # a permuted subset relabels by POSITION in the subset (node_idx[subset] = arange), edge order stays the edge list's
    ei2, ea2 = next_oracle.subgraph([5, 3, 4], edge_index, edge_attr, 7)
    assert np.array_equal(ea2, ea) and np.array_equal(ei2, np.array([[1, 2, 2, 0], [2, 1, 0, 2]]))
