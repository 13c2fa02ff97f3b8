#!/usr/bin/env python
"""Golden fixture for the rows SURVEY.md section 8 marks "next" (f2 pooling / feature hand-off, f4 labels), produced by
running the UNMODIFIED reference functions (build container only: needs /root/reference):

  * src/utils/models_utils.py        EmbeddingProcessor.pool_ngram_embeddings_for_protein_fast   (:210-262)
  * src/pipeline/protgram_directgcn_trainer.py   ProtGramDirectGCNTrainer._generate_next_node_labels (:222-237),
    called unbound on a stand-in `self` that only carries config.DEBUG_VERBOSE (all the method reads)

over the same third-party shim as make_golden.py (+ two unused torch_geometric.utils names the trainer imports).
The feature hand-off (:323-330) sits inline in the trainer's run(): its loop is cut out of the reference module's source at
generation time and executed on stand-in locals (reference_feature_handoff below), which pins oracle/next_oracle.py:init_level_features.

Outputs: tests/golden/next_rows.npz.    Run: python tests/golden/make_golden_next.py
"""
import os
import random
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402


def reference_feature_handoff(ref_tr, node_to_idx, prev_map, prev_emb, n_val):
    """Runs the reference's OWN feature hand-off loop (protgram_directgcn_trainer.py:323-330, inline in run()): its source lines
    are cut out of the module at generation time (from the `for ngram_str, idx in tqdm(` line to the `x[idx] = ...` assignment)
    and executed on stand-in locals -- nothing of it is stored in this repository."""
    import inspect
    import textwrap
    src = inspect.getsource(ref_tr.ProtGramDirectGCNTrainer.run).splitlines()
    a = next(i for i, l in enumerate(src) if "for ngram_str, idx in tqdm(graph_obj.node_to_idx.items()" in l)
    b = next(i for i in range(a, len(src)) if "x[idx] = torch.from_numpy(np.mean(" in src[i])
    loop = textwrap.dedent("\n".join(src[a:b + 1]))
    env = {
        "tqdm": lambda it, **kw: it, "np": np, "torch": torch, "n_val": n_val,
        "graph_obj": types.SimpleNamespace(node_to_idx=node_to_idx),
        "prev_level_ngram_to_idx_map": prev_map, "prev_level_embeds_np": prev_emb,
        "self": types.SimpleNamespace(config=types.SimpleNamespace(DEBUG_VERBOSE=False), device=torch.device("cpu")),
        "x": torch.zeros(len(node_to_idx), prev_emb.shape[1], dtype=torch.float),
    }
    exec(loop, env)
    return env["x"].numpy()


def main():
    ref_db, ref_du, ref_gu, ref_model, ref_mu = mg.import_reference()
    tgu = sys.modules["torch_geometric.utils"]
    tgu.subgraph = tgu.to_networkx = None
    import tempfile
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())
    try:
        from src.pipeline import protgram_directgcn_trainer as ref_tr
    finally:
        os.chdir(cwd)
    from oracle import next_oracle

    rng = np.random.default_rng(7)
    aa = list("ACDEFGHIKLMNPQRSTVWY")
    # proteins: ragged lengths incl. shorter than n, repeats (the distinct-n-gram semantics), an unknown letter
    seqs = []
    for i in range(60):
        L = int(rng.integers(1, 90))
        s = "".join(rng.choice(aa[:6] if i % 3 else aa, size=L))
        if i % 11 == 4:
            s = s[: L // 2] + "X" + s[L // 2:]
        if i % 13 == 6:
            s = "AC" * (L // 2 + 1)
        seqs.append((f"P{i:03d}", s))
    seqs.append(("EMPTY", ""))
    seqs.append(("ONLYX", "XXXX"))
    rec = {"prot_ids": np.array([p for p, _ in seqs]), "prot_seqs": np.array([s for _, s in seqs])}
    fasta = "".join(f">{p}\n{s}\n" for p, s in seqs if s and "X" not in s)
    builds = mg.reference_build(ref_db, ref_du, ref_gu, fasta, 3)
    embs = {}
    for n in (1, 2, 3):
        nodes = [str(x) for x in builds[n]["nodes"]]
        ngram_map = {s: i for i, s in enumerate(nodes)}
        emb = rng.standard_normal((len(nodes), 8)).astype(np.float32)
        embs[n] = (nodes, ngram_map, emb)
        pooled = ref_mu.EmbeddingProcessor.pool_ngram_embeddings_for_protein_fast(seqs, n, ngram_map, emb)
        ids = [p for p, _ in seqs if p in pooled]
        rec[f"n{n}_nodes"] = np.array(nodes)
        rec[f"n{n}_emb"] = emb
        rec[f"n{n}_pooled_ids"] = np.array(ids)
        rec[f"n{n}_pooled"] = np.stack([pooled[p] for p in ids]) if ids else np.zeros((0, 8), np.float32)
        # labels from the reference method (ties: random.choice -> any maximal successor is admissible)
        g = types.SimpleNamespace(number_of_nodes=len(nodes),
                                  A_out_w=torch.sparse_coo_tensor(torch.from_numpy(builds[n]["A_out_w_idx"]),
                                                                  torch.from_numpy(builds[n]["A_out_w_val"]),
                                                                  (len(nodes), len(nodes))).coalesce())
        me = types.SimpleNamespace(config=types.SimpleNamespace(DEBUG_VERBOSE=False))
        random.seed(n)
        labels, ncls = ref_tr.ProtGramDirectGCNTrainer._generate_next_node_labels(me, g)
        assert ncls == len(nodes)
        rec[f"n{n}_labels_ref"] = labels.numpy()
        rec[f"n{n}_a_out_idx"] = builds[n]["A_out_w_idx"]
        rec[f"n{n}_a_out_val"] = builds[n]["A_out_w_val"]
        if n > 1:
            prev_nodes, prev_map, prev_emb = embs[n - 1]
            rec[f"n{n}_x_init_oracle"] = next_oracle.init_level_features(nodes, prev_map, prev_emb)
            rec[f"n{n}_x_init_ref"] = reference_feature_handoff(ref_tr, ngram_map, prev_map, prev_emb, n)
            assert np.array_equal(rec[f"n{n}_x_init_ref"], rec[f"n{n}_x_init_oracle"])   # the oracle is pinned by the reference's own loop
    np.savez_compressed(os.path.join(HERE, "next_rows.npz"), **rec)
    print("wrote next_rows.npz", {k: v.shape for k, v in rec.items() if k.endswith("pooled")})


if __name__ == "__main__":
    main()
