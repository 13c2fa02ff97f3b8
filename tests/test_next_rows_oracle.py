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
