#!/usr/bin/env python
"""Generate the committed golden fixtures by running the UNMODIFIED reference.

This script only runs in the build container (it needs /root/reference, which
does not exist on the GPU box).  It imports the reference's own modules
verbatim:

  * src/pipeline/data_builder.py   (_preprocess..., _extract_ngrams..., _extract_edges...)
  * src/utils/data_utils.py        (DataLoader.parse_sequences)
  * src/utils/graph_utils.py       (DirectedNgramGraph)
  * src/models/protgram_directgcn.py (DirectGCNLayer, ProtGramDirectGCN)
  * src/utils/models_utils.py      (EmbeddingProcessor.l2_normalize_torch)

over a ~40 line shim for the third-party packages that are absent here
(torch_geometric, dask, h5py, Bio, requests).  The shim restates the *published*
semantics of the PyG pieces the reference calls:

  MessagePassing(aggr='add').propagate(ei, x=, edge_weight=)
        -> out = zeros(N, F); out.index_add_(0, ei[1], message(x[ei[0]], w))
  add_self_loops(ei, num_nodes) -> appends arange(N) twice, unconditionally
  degree(index, N, dtype)       -> occurrence count of each index

Dask is only used by GraphBuilder.run() for `distinct()` (a set) and
`groupby(['source','target']).size()` (exact integer counts); the driver below
reproduces run()'s data flow (data_builder.py:97-105,151-177,203-220,267-311)
with a Python set / collections.Counter and the reference's own helper
functions, writes the aggregated-edges parquet exactly like
data_builder.py:281-286 and then lets the reference's DirectedNgramGraph read it.

Outputs: tests/golden/*.npz  (small; committed).
Run:     python tests/golden/make_golden.py
"""
import collections
import os
import sys
import tempfile
import types

import numpy as np
import pandas as pd
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


# ----------------------------------------------------------------------------
# third-party shim
# ----------------------------------------------------------------------------
def install_shim():
    def mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m

    for name in ("dask", "dask.bag", "dask.dataframe", "h5py", "Bio", "requests"):
        mod(name)
    sys.modules["Bio"].SeqIO = None
    sys.modules["dask"].bag = sys.modules["dask.bag"]
    sys.modules["dask"].dataframe = sys.modules["dask.dataframe"]

    tg = mod("torch_geometric")
    tg_data = mod("torch_geometric.data")
    tg_nn = mod("torch_geometric.nn")
    tg_utils = mod("torch_geometric.utils")
    tg.data, tg.nn, tg.utils = tg_data, tg_nn, tg_utils

    class Data:
        def __init__(self, **kw):
            for k, v in kw.items():
                setattr(self, k, v)

        def to(self, device):
            for k, v in list(self.__dict__.items()):
                if torch.is_tensor(v):
                    setattr(self, k, v.to(device))
            return self

    class MessagePassing(torch.nn.Module):
        def __init__(self, aggr="add"):
            super().__init__()
            assert aggr == "add"

        def propagate(self, edge_index, x, edge_weight=None):
            msg = self.message(x[edge_index[0]], edge_weight)
            out = torch.zeros(x.size(0), x.size(1), dtype=x.dtype, device=x.device)
            return out.index_add_(0, edge_index[1], msg)

    def add_self_loops(edge_index, edge_attr=None, fill_value=None, num_nodes=None):
        loop = torch.arange(num_nodes, dtype=edge_index.dtype, device=edge_index.device)
        return torch.cat([edge_index, loop.unsqueeze(0).repeat(2, 1)], dim=1), edge_attr

    def degree(index, num_nodes=None, dtype=None):
        out = torch.zeros(num_nodes, dtype=dtype, device=index.device)
        return out.scatter_add_(0, index, torch.ones(index.numel(), dtype=dtype, device=index.device))

    tg_data.Data = Data
    tg_nn.MessagePassing = MessagePassing
    tg_utils.add_self_loops = add_self_loops
    tg_utils.degree = degree


def import_reference():
    install_shim()
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())  # reference Config() resolves Path(".")
    try:
        from src.pipeline import data_builder as ref_db
        from src.utils import data_utils as ref_du
        from src.utils import graph_utils as ref_gu
        from src.models import protgram_directgcn as ref_model
        from src.utils import models_utils as ref_mu
    finally:
        os.chdir(cwd)
    return ref_db, ref_du, ref_gu, ref_model, ref_mu


# ----------------------------------------------------------------------------
# builder goldens
# ----------------------------------------------------------------------------
def reference_build(ref_db, ref_du, ref_gu, fasta_text, n_max, eps=1e-9):
    """GraphBuilder.run() data flow with the reference's own helpers."""
    tmp = tempfile.mkdtemp()
    fasta = os.path.join(tmp, "in.fasta")
    with open(fasta, "w") as f:
        f.write(fasta_text)
    stream = []
    first = True
    for tup in ref_du.DataLoader.parse_sequences(fasta):  # data_builder.py:97-102
        stream.append((tup, first))
        first = False
    pre = [ref_db._preprocess_sequence_tuple_for_bag(t, flag) for t, flag in stream]
    out = {}
    for n in range(1, n_max + 1):
        uniq = set()
        for t in pre:
            uniq.update(ref_db._extract_ngrams_from_sequence_tuple(t, n))
        df = pd.DataFrame(sorted(uniq), columns=["ngram"])  # :164
        df = df.sort_values("ngram").reset_index(drop=True)  # :172
        df["id"] = df.index
        ngram_to_id = df.set_index("ngram")["id"].to_dict()
        cnt = collections.Counter()
        for t in pre:
            for line in ref_db._extract_edges_from_sequence_tuple(t, n, ngram_to_id):
                s, d = line.split()
                cnt[(int(s), int(d))] += 1
        idx_to_node = df.set_index("id")["ngram"].to_dict()
        edge_file = os.path.join(tmp, f"agg_n{n}.parquet")
        if cnt:
            # dd.groupby(['source','target']).size() -> frame(source,target,weight)
            keys = sorted(cnt)
            edf = pd.DataFrame({"source": [k[0] for k in keys], "target": [k[1] for k in keys],
                                "weight": [cnt[k] for k in keys]})
            # shuffle rows: groupby output order is unspecified; the graph class must not care
            edf = edf.sample(frac=1.0, random_state=n).reset_index(drop=True)
            edf.to_parquet(edge_file, index=False)
        g = ref_gu.DirectedNgramGraph(nodes=idx_to_node, edge_file_path=edge_file,
                                      epsilon_propagation=eps, n_value=n)
        rec = {"nodes": np.array([idx_to_node[i] for i in range(len(idx_to_node))], dtype=object).astype("U"),
               "n_transitions": np.int64(sum(cnt.values())),
               "number_of_nodes": np.int64(g.number_of_nodes),
               "number_of_edges": np.int64(g.number_of_edges)}
        for name in ("A_out_w", "A_in_w", "A_undirected_norm_sparse", "mathcal_A_out", "mathcal_A_in"):
            t = getattr(g, name).coalesce()
            rec[name + "_idx"] = t.indices().numpy().astype(np.int64)
            rec[name + "_val"] = t.values().numpy().astype(np.float32)
        out[n] = rec
    return out


def synth_fasta(rng, nseq, lmin, lmax, alphabet, weird=False):
    lines = []
    for i in range(nseq):
        L = int(rng.integers(lmin, lmax + 1))
        s = "".join(rng.choice(list(alphabet), size=L))
        if weird and i % 17 == 3:
            s = s.lower()  # parse_sequences upper-cases (data_utils.py:207)
        hdr = f">sp|P{i:05d}|X" if i % 2 == 0 else f">id{i} desc"
        if weird and i % 23 == 5 and L > 10:
            # multi-line record + blank line in the middle
            lines += [hdr, s[:7], "", s[7:]]
        else:
            lines += [hdr, s]
    return "\n".join(lines) + "\n"


def save_build(name, fasta_text, recs):
    flat = {"fasta": np.array(fasta_text)}
    for n, rec in recs.items():
        for k, v in rec.items():
            flat[f"n{n}_{k}"] = v
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **flat)
    print("wrote", name, {n: (int(r["number_of_nodes"]), int(r["number_of_edges"])) for n, r in recs.items()})


# ----------------------------------------------------------------------------
# model goldens
# ----------------------------------------------------------------------------
def random_edges(rng, N, E, weighted):
    src = rng.integers(0, N, size=E)
    dst = rng.integers(0, N, size=E)
    ei = torch.from_numpy(np.stack([src, dst]).astype(np.int64))
    ew = torch.from_numpy(rng.random(E).astype(np.float32) + 0.1) if weighted else None
    return ei, ew


def randomise_(module, gen):
    """Non-trivial parameter values (reference init leaves biases 0 and gates 1)."""
    with torch.no_grad():
        for name, p in module.named_parameters():
            if p.ndim >= 2 and "constant" not in name and "C_" not in name:
                continue  # keep xavier weights
            p.add_(0.25 * torch.randn(p.shape, generator=gen))


def model_golden(ref_model, name, N, dims, edges, use_vec, n_gram_len, one_gram_dim, num_classes,
                 original_indices=None, num_graph_nodes=None, seed=0):
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    num_graph_nodes = N if num_graph_nodes is None else num_graph_nodes
    model = ref_model.ProtGramDirectGCN(layer_dims=dims, num_graph_nodes=num_graph_nodes,
                                        task_num_output_classes=num_classes, n_gram_len=n_gram_len,
                                        one_gram_dim=one_gram_dim, max_pe_len=16, dropout=0.0,
                                        use_vector_coeffs=use_vec)
    randomise_(model, gen)
    x = torch.randn(N, dims[0], generator=gen)
    (ei_in, ew_in), (ei_out, ew_out), (ei_un, ew_un) = edges
    from torch_geometric.data import Data
    data = Data(x=x.clone().requires_grad_(True), edge_index_in=ei_in, edge_weight_in=ew_in,
                edge_index_out=ei_out, edge_weight_out=ew_out,
                edge_index_undirected_norm=ei_un, edge_weight_undirected_norm=ew_un)
    if original_indices is not None:
        data.original_indices = original_indices
    # per-layer activations (protgram_directgcn.py:210-216) recorded with hooks
    layer_out = []
    hooks = [c.register_forward_hook(lambda m, i, o: layer_out.append(o.detach().clone())) for c in model.convs]
    model.train()  # dropout=0.0 in the stack; decoder Dropout(0.5) is active -> use eval for logits
    model.eval()
    logp, emb = model(data=data)
    for h in hooks:
        h.remove()
    y = torch.randint(0, num_classes, (N,), generator=gen)
    wvec = torch.randn(emb.shape, generator=gen)
    loss = torch.nn.functional.nll_loss(logp, y) + (emb * wvec).sum()
    loss.backward()
    rec = {"x": x.numpy(), "y": y.numpy(), "wvec": wvec.numpy(),
           "logp": logp.detach().numpy(), "emb": emb.detach().numpy(),
           "loss": np.float32(loss.item()), "grad_x": data.x.grad.numpy(),
           "dims": np.array(dims), "use_vec": np.bool_(use_vec), "n_gram_len": np.int64(n_gram_len),
           "one_gram_dim": np.int64(one_gram_dim), "num_classes": np.int64(num_classes),
           "num_graph_nodes": np.int64(num_graph_nodes)}
    for i, lo in enumerate(layer_out):
        rec[f"layer{i}_out"] = lo.numpy()
    for k, (ei, ew) in zip(("in", "out", "und"), edges):
        rec[f"ei_{k}"] = ei.numpy()
        if ew is not None:
            rec[f"ew_{k}"] = ew.numpy()
    if original_indices is not None:
        rec["original_indices"] = original_indices.numpy()
    for k, v in model.state_dict().items():
        rec["sd:" + k] = v.numpy()
    for k, p in model.named_parameters():
        rec["grad:" + k] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
    print("wrote", name, "loss", float(loss))


def main():
    ref_db, ref_du, ref_gu, ref_model, ref_mu = import_reference()
    rng = np.random.default_rng(42)

    # KA1: run_graph_builder.py:24-28, n<=3.  KA2: unit_tests.py:41, n=1.
    ka1 = ">seq1\nACGTACT\n>seq2\nTTACGTT\n>seq3\nAGATAGA\n"
    ka2 = ">seq1\nACGT\n>seq2\nTTAC\n>seq3\nAGA\n"
    save_build("build_ka1", ka1, reference_build(ref_db, ref_du, ref_gu, ka1, 3))
    save_build("build_ka2", ka2, reference_build(ref_db, ref_du, ref_gu, ka2, 1))

    # ragged / tiny / weird records: lengths 1..40 (sequences shorter than n), lower case,
    # multi-line records, the 20 amino acids plus X/B/Z/U/O/* (data-defined alphabet).
    aa = "ACDEFGHIKLMNPQRSTVWY"
    rag = synth_fasta(rng, 120, 1, 40, aa + "XBZUO*", weird=True)
    save_build("build_ragged", rag, reference_build(ref_db, ref_du, ref_gu, rag, 4))

    # protein-like: 300 x ~120 residues, 20-letter alphabet, n<=3
    prot = synth_fasta(rng, 300, 80, 160, aa)
    save_build("build_protein", prot, reference_build(ref_db, ref_du, ref_gu, prot, 3))

    # low-complexity: homopolymer runs, single-residue sequences, first sequence of length 1
    low = ">a\nA\n>b\nAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA\n>c\nC\n>d\nACACACACACACACACACACACAC\n>e\nAA\n"
    save_build("build_lowcomplexity", low, reference_build(ref_db, ref_du, ref_gu, low, 3))

    # ---- model goldens ----
    # (1) reference-built graph (shared symmetric pattern), vector gates, n=1 style PE
    recs = reference_build(ref_db, ref_du, ref_gu, prot, 2)
    r = recs[2]
    N = int(r["number_of_nodes"])
    edges = tuple((torch.from_numpy(r[k + "_idx"]), torch.from_numpy(r[k + "_val"]))
                  for k in ("mathcal_A_in", "mathcal_A_out", "A_undirected_norm_sparse"))
    model_golden(ref_model, "model_refgraph", N, [24, 40, 16, 8], edges, True, 2, 0, 7, seed=1)

    # (2) n=1 graph with positional encoding active (x width == n_gram_len*one_gram_dim)
    r1 = recs[1]
    N1 = int(r1["number_of_nodes"])
    edges1 = tuple((torch.from_numpy(r1[k + "_idx"]), torch.from_numpy(r1[k + "_val"]))
                   for k in ("mathcal_A_in", "mathcal_A_out", "A_undirected_norm_sparse"))
    model_golden(ref_model, "model_n1_pe", N1, [32, 16, 16], edges1, True, 1, 32, N1, seed=2)

    # (3) benchmarker-style: unsymmetric, unweighted, duplicate edges (gnn_benchmarker.py:297-305),
    #     scalar gates (use_vector_coeffs=False)
    Nb = 57
    ei_out, _ = random_edges(rng, Nb, 400, False)
    ei_in = ei_out[[1, 0]].contiguous()
    ei_un, ew_un = random_edges(rng, Nb, 500, True)
    model_golden(ref_model, "model_general_scalar", Nb, [12, 20, 20, 5],
                 ((ei_in, None), (ei_out, None), (ei_un, ew_un)), False, 3, 0, 4, seed=3)

    # (4) cluster mini-batch style: original_indices gather of per-node gates / constant
    #     (protgram_directgcn.py:116-120): sub-graph of 30 nodes out of a 90-node parent
    Ns, Np = 30, 90
    oi = torch.from_numpy(rng.choice(Np, size=Ns, replace=False).astype(np.int64))
    e_in = random_edges(rng, Ns, 150, True)
    e_out = random_edges(rng, Ns, 150, True)
    e_un = random_edges(rng, Ns, 200, True)
    model_golden(ref_model, "model_cluster_batch", Ns, [10, 12, 6], (e_in, e_out, e_un), True, 2, 0, 3,
                 original_indices=oi, num_graph_nodes=Np, seed=4)


if __name__ == "__main__":
    main()
