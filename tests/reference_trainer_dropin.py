#!/usr/bin/env python
"""Drop-in check with the reference's OWN trainer code (run as a subprocess by tests/test_reference_driver_dropin.py, build
container only: needs /root/reference).  The central block of ProtGramDirectGCNTrainer.run() -- load the graph, move the raw
adjacency to the device and re-create the propagation matrices (:290-299), `_generate_next_node_labels` (:222-237), the Data
object (:347-367), `_train_model_full_batch` (:76-108) and `EmbeddingProcessor.extract_gcn_node_embeddings`
(models_utils.py:265-273) -- is executed twice by the reference's unmodified functions:
  (R) on the reference's own DirectedNgramGraph / ProtGramDirectGCN (third-party packages shimmed as in tests/golden/make_golden.py)
  (M) on this package's same-named classes (native entry points replaced by their executable spec: no GPU here)
with identical initial parameters, seeds and optimizer, and the outcomes are printed as JSON: label equality, final training
loss of both, max differences of the trained parameters and of the extracted embeddings."""
import io
import json
import os
import random
import re
import sys
import tempfile
from contextlib import redirect_stdout

import numpy as np
import pandas as pd
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def main():
    import make_golden as mg
    ref_db, ref_du, ref_gu, ref_model, ref_mu = mg.import_reference()
    tgu = sys.modules["torch_geometric.utils"]
    tgu.subgraph = tgu.to_networkx = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("not on this path"))
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())
    try:
        from src.pipeline import protgram_directgcn_trainer as ref_tr
        from config import Config
        cfg = Config()
    finally:
        os.chdir(cwd)
    cfg.DEBUG_VERBOSE = True
    trainer = ref_tr.ProtGramDirectGCNTrainer(cfg)
    trainer.device = torch.device("cpu")

    import protgram_directgcn_b200 as pg
    from protgram_directgcn_b200 import _native as nat
    from tests import kernel_spec
    kernel_spec.install_plain(nat)

    rng = np.random.default_rng(7)
    fasta = mg.synth_fasta(rng, 80, 15, 60, "ACDEFGHIKLMNPQRSTVWY")
    n = 2
    rec = mg.reference_build(ref_db, ref_du, ref_gu, fasta, n)[n]
    tmp = tempfile.mkdtemp()
    idx, val = rec["A_out_w_idx"], rec["A_out_w_val"]
    edge_file = os.path.join(tmp, "edges.parquet")
    pd.DataFrame({"source": idx[0], "target": idx[1], "weight": val.astype(np.int64)}).to_parquet(edge_file, index=False)
    nodes = {i: str(s) for i, s in enumerate(rec["nodes"])}
    dims, epochs, lr, l2 = [16, 32, 16], 6, 0.05, 1e-4
    x0 = torch.randn(len(nodes), dims[0], generator=torch.Generator().manual_seed(3))
    init = None
    out = {}
    for tag, graph_cls, model_cls, save, load in (("R", ref_gu.DirectedNgramGraph, ref_model.ProtGramDirectGCN, ref_du.DataUtils.save_object,
                                                   ref_du.DataUtils.load_object),
                                                  ("M", pg.DirectedNgramGraph, pg.ProtGramDirectGCN, pg.DataUtils.save_object,
                                                   pg.DataUtils.load_object)):
        graph = graph_cls(nodes=nodes, edge_file_path=edge_file, epsilon_propagation=1e-9, n_value=n)
        path = os.path.join(tmp, f"graph_{tag}.pkl")
        save(graph, path)
        graph_obj = load(path)                                          # trainer :289
        graph_obj.n_value = n
        graph_obj.A_out_w = graph_obj.A_out_w.to(trainer.device)        # :295-297
        graph_obj.A_in_w = graph_obj.A_in_w.to(trainer.device)
        graph_obj.A_undirected_norm_sparse = graph_obj.A_undirected_norm_sparse.to(trainer.device)
        graph_obj._create_propagation_matrices_for_gcn()                # :299
        random.seed(0)
        with redirect_stdout(io.StringIO()):
            labels, num_classes = trainer._generate_next_node_labels(graph_obj)          # the reference's function on either graph
        Data = sys.modules["torch_geometric.data"].Data
        full_data = Data(x=x0.clone(), y=labels)
        full_data.num_nodes = graph_obj.number_of_nodes
        model = model_cls(layer_dims=dims, num_graph_nodes=graph_obj.number_of_nodes, task_num_output_classes=num_classes,
                          n_gram_len=n, one_gram_dim=0, max_pe_len=cfg.GCN_MAX_PE_LEN, dropout=0.0, use_vector_coeffs=True)
        if init is None:
            init = {k: v.clone() for k, v in model.state_dict().items()}
        assert set(init) == set(model.state_dict())
        model.load_state_dict(init, strict=True)
        optimizer = torch.optim.SGD(model.parameters(), lr=lr)
        full_data.edge_index_in = graph_obj.mathcal_A_in.indices()      # :362-367
        full_data.edge_weight_in = graph_obj.mathcal_A_in.values()
        full_data.edge_index_out = graph_obj.mathcal_A_out.indices()
        full_data.edge_weight_out = graph_obj.mathcal_A_out.values()
        full_data.edge_index_undirected_norm = graph_obj.A_undirected_norm_sparse.indices()
        full_data.edge_weight_undirected_norm = graph_obj.A_undirected_norm_sparse.values()
        torch.manual_seed(11)                                           # the decoder's Dropout(0.5) draws from the global generator
        log = io.StringIO()
        with redirect_stdout(log):
            trainer._train_model_full_batch(model, full_data, optimizer, epochs, "next_node", l2)
        m = re.search(r"Total Loss: ([0-9.]+), Primary Loss: ([0-9.]+)", log.getvalue())
        emb = ref_mu.EmbeddingProcessor.extract_gcn_node_embeddings(model, full_data, trainer.device)
        out[tag] = {"labels": labels.numpy(), "loss": float(m.group(1)), "primary": float(m.group(2)),
                    "params": {k: v.detach().numpy().copy() for k, v in model.state_dict().items()}, "emb": emb}
    r, m = out["R"], out["M"]
    moved = max(float(np.max(np.abs(r["params"][k] - init[k].numpy()))) for k in init)
    print(json.dumps({
        "nodes": len(nodes), "classes": int(num_classes), "epochs": epochs,
        "labels_equal": bool(np.array_equal(r["labels"], m["labels"])),
        "loss_ref": r["loss"], "loss_mine": m["loss"],
        "param_max_abs_diff": max(float(np.max(np.abs(r["params"][k] - m["params"][k]))) for k in init),
        "param_max_move": moved,
        "emb_max_abs_diff": float(np.max(np.abs(r["emb"] - m["emb"]))), "emb_max_abs": float(np.max(np.abs(r["emb"]))),
    }))


def full_run():
    """The reference's whole `ProtGramDirectGCNTrainer.run()` (graph loading, per-level feature hand-off, label generation,
    training with the config's Adam / scheduler / early stopping, embedding extraction, protein pooling, H5 writing) executed
    unmodified twice: with the reference's classes on the reference-built pickles, and with `ProtGramDirectGCN`,
    `DirectedNgramGraph`, `DataUtils` of the trainer module swapped for this package's on pickles written by this package's
    GraphBuilder.  Prints the max difference of the pooled protein embeddings it "saves" (h5py is a recording stub)."""
    import make_golden as mg
    ref_db, ref_du, ref_gu, ref_model, ref_mu = mg.import_reference()
    tgu = sys.modules["torch_geometric.utils"]
    tgu.subgraph = tgu.to_networkx = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("not on this path"))
    saved = {}

    class _H5File:
        def __init__(self, path, mode="r"):
            self.store = saved.setdefault(os.path.basename(path), {})

        def __enter__(self):
            return self

        def __exit__(self, *exc):
            return False

        def create_dataset(self, key, data=None):
            self.store[key] = np.array(data)

    sys.modules["h5py"].File = _H5File
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())
    try:
        from src.pipeline import protgram_directgcn_trainer as ref_tr
        from config import Config
    finally:
        os.chdir(cwd)
    import protgram_directgcn_b200 as pg
    from protgram_directgcn_b200 import _native as nat
    from tests import kernel_spec
    kernel_spec.install_plain(nat)

    rng = np.random.default_rng(21)
    fasta_text = mg.synth_fasta(rng, 60, 12, 50, "ACDEFGHIKLMNPQRSTVWY", weird=True)
    n_max = 2
    results = {}
    for tag in ("R", "M"):
        tmp = tempfile.mkdtemp()
        fasta = os.path.join(tmp, "in.fasta")
        with open(fasta, "w") as fh:
            fh.write(fasta_text)
        os.chdir(tmp)
        try:
            cfg = Config()
        finally:
            os.chdir(cwd)
        from pathlib import Path
        cfg.GCN_INPUT_FASTA_PATH = Path(fasta)
        cfg.BASE_OUTPUT_DIR = Path(tmp) / "out"
        cfg.GRAPH_OBJECTS_DIR = cfg.BASE_OUTPUT_DIR / "1_graph_objects"
        cfg.GCN_EMBEDDINGS_DIR = cfg.BASE_OUTPUT_DIR / "2_gcn_embeddings"
        cfg.GCN_NGRAM_MAX_N = n_max
        cfg.GCN_HIDDEN_LAYER_DIMS = [24, 16]
        cfg.GCN_1GRAM_INIT_DIM = 12
        cfg.GCN_EPOCHS_PER_LEVEL = 4
        cfg.GCN_TASK_TYPES_PER_LEVEL = {1: "next_node", 2: "next_node"}
        cfg.GCN_USE_CLUSTER_TRAINING = False
        # the reference's own _apply_pe (protgram_directgcn.py:185-191) raises "a leaf Variable that requires grad is being used in an
        # in-place operation" under this torch (2.11) when x carries no grad, i.e. in its own n=1 training: positional encoding off
        cfg.GCN_MAX_PE_LEN = 0
        cfg.ID_MAPPING_MODE = "none"
        cfg.APPLY_PCA_TO_GCN = False
        cfg.GCN_RUN_SANITY_CHECK_PPI = False
        cfg.DEBUG_VERBOSE = False
        os.makedirs(cfg.GRAPH_OBJECTS_DIR, exist_ok=True)
        if tag == "R":
            # reference pickles: run()'s data flow with the reference's helpers (Dask replaced by set / Counter, see make_golden.py)
            recs = mg.reference_build(ref_db, ref_du, ref_gu, fasta_text, n_max)
            for n, rec in recs.items():
                edge_file = os.path.join(tmp, f"e{n}.parquet")
                pd.DataFrame({"source": rec["A_out_w_idx"][0], "target": rec["A_out_w_idx"][1],
                              "weight": rec["A_out_w_val"].astype(np.int64)}).to_parquet(edge_file, index=False)
                g = ref_gu.DirectedNgramGraph(nodes={i: str(s) for i, s in enumerate(rec["nodes"])}, edge_file_path=edge_file,
                                              epsilon_propagation=1e-9, n_value=n)
                ref_du.DataUtils.save_object(g, str(cfg.GRAPH_OBJECTS_DIR / f"ngram_graph_n{n}.pkl"))
            ref_tr.ProtGramDirectGCN, ref_tr.DirectedNgramGraph, ref_tr.DataUtils = ref_model.ProtGramDirectGCN, ref_gu.DirectedNgramGraph, ref_du.DataUtils
        else:
            with redirect_stdout(io.StringIO()):
                pg.GraphBuilder(cfg).run()          # the reference Config object drives this package's builder
            ref_tr.ProtGramDirectGCN, ref_tr.DirectedNgramGraph, ref_tr.DataUtils = pg.ProtGramDirectGCN, pg.DirectedNgramGraph, pg.DataUtils
        trainer = ref_tr.ProtGramDirectGCNTrainer(cfg)
        trainer.device = torch.device("cpu")
        torch.manual_seed(5)
        random.seed(5)
        np.random.seed(5)
        saved.clear()
        log = io.StringIO()
        with redirect_stdout(log):
            trainer.run()
        assert "SUCCESS: Primary embeddings saved" in log.getvalue(), log.getvalue()[-2000:]
        (fname, store), = saved.items()
        results[tag] = (fname, {k: v.copy() for k, v in store.items()})
    (fr, r), (fm, m) = results["R"], results["M"]
    keys_equal = fr == fm and sorted(r) == sorted(m)
    diff = max(float(np.max(np.abs(r[k].astype(np.float64) - m[k].astype(np.float64)))) for k in r) if keys_equal else float("inf")
    scale = max(float(np.max(np.abs(v))) for v in r.values())
    print(json.dumps({"mode": "full_run", "file": fr, "proteins": len(r), "keys_equal": bool(keys_equal), "dim": int(next(iter(r.values())).shape[0]),
                      "pooled_max_abs_diff": diff, "pooled_max_abs": scale}))


if __name__ == "__main__":
    full_run() if "--full-run" in sys.argv else main()
