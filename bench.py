#!/usr/bin/env python
"""bench.py -- headline benchmark of the two hot paths on BASELINE.json config C2:

    "n=3 graph (8k nodes) from 500k synthetic 350-residue sequences, DirectGCN train +
     embedding extraction on 1xB200"

One STEP = one pass of the hot path over one corpus batch:
    A  alphabet discovery -> (n+1)-gram count -> [NCCL sum over ranks] -> node ids / edge table
       -> adjacency + propagation matrices                          (libpgb200: ngram.cu, graph.cu)
    B  one DirectGCN training step (fwd, nll + L2 term, bwd, Adam) on the graph just built,
       dims 64->256->128->64, then the eval-mode embedding extraction   (spmm.cu, gemm.cu)

metric = residues/s through the whole step (whole job, all ranks).  `value` has the corpus
already resident in HBM; `e2e` runs the same step from PINNED HOST bytes through the
reference-facing classes (H2D of the corpus, graph object materialised on the host like the
reference's (D2H of all five matrices, node names decoded), the DirectGCN step fed from the device-side CSR
the builder keeps next to that object, embeddings read back to numpy); every e2e step issues one
full H2D copy of the corpus, double-buffered on a copy stream so that the upload of batch i+1
overlaps the graph/DirectGCN kernels of batch i (host/corpus.py:CorpusUploader).

    python bench.py [--gpus N --steps K --warmup W]            our arm   (torchrun for N > 1)
    python bench.py --impl reference [...]                     the reference algorithm on host cores

The reference is pure Python (nothing compiles, nothing pip-installs without PyG/Dask), so the
reference arm times the in-repo CPU restatement of its algorithm (oracle/, kind="port") with
all host cores on a bounded sample -- see cpu_baseline.sample in the JSON line.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NSEQ, SEQ_LEN, N_LEVEL, SEED = 500_000, 350, 3, 42
DIMS = [64, 256, 128, 64]
L2_LAMBDA, LR, DROPOUT = 1e-7, 1e-3, 0.5
METRIC = "c2_pipeline_residues_per_s"
WORKLOAD = ("C2: n=3 transition graph from 500k x 350-residue synthetic sequences per GPU + one DirectGCN "
            "train step (64-256-128-64, next_node task) + embedding extraction")


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# --------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# --------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
class B200Pipeline:
    def __init__(self, rank, world, device):
        import protgram_directgcn_b200 as pg
        from protgram_directgcn_b200 import _native as nat
        from protgram_directgcn_b200.host import corpus, data_builder, graph_utils
        self.pg, self.nat, self.corpus, self.db, self.gu = pg, nat, corpus, data_builder, graph_utils
        self.rank, self.world, self.dev = rank, world, device
        self.use_cuda_graph = True
        self.group = None
        if world > 1:
            import torch.distributed as dist
            self.group = dist.group.WORLD
        self.nbytes = NSEQ * (SEQ_LEN + 2) + (1 if rank == 0 else 0)
        self.d_buf = torch.empty(self.nbytes, dtype=torch.uint8, device=device)
        nat.call("pg_synth_corpus", nat.ptr(self.d_buf), rank * NSEQ, NSEQ, SEQ_LEN, SEED, int(rank == 0), nat.stream_ptr())
        self.h_buf = None
        self.up = None
        self.count_ms = []
        self.ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        self.model = self.opt = self.x = self.labels = self.graphed = None
        # Two-stage software pipeline over consecutive batches: the graph build of batch k+1 (its own stream; its three small
        # host reads -- alphabet, node / edge counts, pattern size -- only wait for THAT stream) runs under the DirectGCN replay
        # of batch k on the main stream.  pipelined = False restores the strictly sequential step.
        self.pipelined = True
        self.build_stream = torch.cuda.Stream(device=device)
        self.pending = None
        self._build_stream_primed = False
        self.host_bytes = False
        self.h_packed = None

    # ---- hot path A on a device-resident corpus; returns the device edge table + matrices
    def build(self, d_buf, materialise_host: bool, after_count=None):
        nat, db = self.nat, self.db
        symbols, d_rank = self.corpus.discover_alphabet(d_buf, self.group)
        sigma = int(symbols.size)
        pow_n, pow_m = db.table_sizes(N_LEVEL, sigma)
        bins = torch.zeros(pow_m, dtype=torch.int64, device=self.dev)
        short = torch.zeros(pow_n, dtype=torch.uint8, device=self.dev)
        self.ev[0].record()
        db.count_level(d_buf, N_LEVEL, d_rank, sigma, bins, short)
        self.ev[1].record()
        self._pending_count_event = True
        if after_count is not None:
            after_count()          # d_buf has no reader left: the uploader may reuse the slot / prefetch the next batch
        if self.group is not None:
            import torch.distributed as dist
            dist.all_reduce(bins, op=dist.ReduceOp.SUM, group=self.group)
            s32 = short.to(torch.int32)
            dist.all_reduce(s32, op=dist.ReduceOp.MAX, group=self.group)
            short = s32.to(torch.uint8)
        node_code, src, dst, cnt = db.extract_level(bins, short, N_LEVEL, sigma)
        names = self.corpus.LazyNodeNames(node_code, symbols, N_LEVEL)   # node id -> n-gram string: async D2H now, decoded in finish()
        graph = self.gu.DirectedNgramGraph.from_edge_arrays(names, src, dst, cnt.to(torch.float32), n_value=N_LEVEL,
                                                            assume_coalesced=True,
                                                            result_device="cpu" if materialise_host else self.dev)
        return graph

    def ensure_model(self, graph):
        if self.model is not None:
            return
        torch.manual_seed(SEED)
        n = graph.number_of_nodes
        self.num_nodes = n
        self.model = self.pg.ProtGramDirectGCN(DIMS, n, n, N_LEVEL, 0, 512, DROPOUT, True).to(self.dev)
        self.opt = torch.optim.Adam(self.model.parameters(), lr=LR, fused=True, capturable=True)
        self.params = [p for p in self.model.parameters() if p.requires_grad]
        g = torch.Generator().manual_seed(SEED)
        self.x = torch.randn(n, DIMS[0], generator=g).to(self.dev)
        self.labels = self.pg.generate_next_node_labels(graph)[0].to(self.dev)    # row f4 kernel (deterministic: first maximal successor)
        self.graphed = None
        if self.use_cuda_graph:
            from protgram_directgcn_b200.host.graphed_step import GraphedDirectGCNStep
            cap = int(graph.mathcal_A_out._nnz() * 1.05) + 1024
            self.graphed = GraphedDirectGCNStep(self.model, self.opt, self.x, self.labels, n, cap, l2_lambda=L2_LAMBDA)

    def train_and_extract(self, graph):
        self.ensure_model(graph)
        if self.graphed is not None:
            side = getattr(graph, "_pg_device", None)
            if side is None or side["col"].device != self.dev:   # graph object came back from the host: trainer-style .to(device)
                m_in, m_out, m_un = graph.mathcal_A_in.to(self.dev), graph.mathcal_A_out.to(self.dev), graph.A_undirected_norm_sparse.to(self.dev)
                rowptr = torch.zeros(graph.number_of_nodes + 1, dtype=torch.int64, device=self.dev)
                rows = m_in.indices()[0].contiguous()
                self.nat.call("pg_rowptr_from_sorted", self.nat.ptr(rows), rows.numel(), graph.number_of_nodes, self.nat.ptr(rowptr), self.nat.stream_ptr())
                self.graphed.load_structure(rowptr, m_in.indices()[1].to(torch.int32), m_in.values(), m_out.values(), m_un.values())
            else:
                self.graphed.load_structure(side["rowptr"], side["col"], side["val_in"], side["val_out"], side["val_und"])
            return self.graphed.replay()
        data = graph.gcn_data(self.x, self.dev)
        self.model.train()
        self.opt.zero_grad(set_to_none=True)
        logp, _ = self.model(data=data)
        loss = torch.nn.functional.nll_loss(logp, self.labels)
        loss.backward()
        # L2 term of the trainer (protgram_directgcn_trainer.py:96-97): loss += lambda * sum ||p||^2.
        # Same value and same gradient (2*lambda*p), computed with multi-tensor ops instead of ~60 x 3 tiny kernels.
        l2 = torch.stack(torch._foreach_norm(self.params)).square().sum()
        torch._foreach_add_([p.grad for p in self.params], self.params, alpha=2.0 * L2_LAMBDA)
        loss = loss.detach() + L2_LAMBDA * l2
        self.opt.step()
        self.model.eval()
        with torch.no_grad():
            _, emb = self.model(data=data)
        return loss, emb

    def _build_on_side_stream(self, get_buf, **kw):
        """build() on the build stream; the main stream then waits for it (and takes co-ownership of the device CSR)."""
        main = torch.cuda.current_stream(self.dev)
        if not self.pipelined:
            return self.build(get_buf(), **kw)
        if not self._build_stream_primed:          # the corpus buffer was generated on the main stream
            self.build_stream.wait_stream(main)
            self._build_stream_primed = True
        with torch.cuda.stream(self.build_stream):
            graph = self.build(get_buf(), **kw)
            ready = torch.cuda.Event()
            ready.record()
        main.wait_event(ready)
        for t in (getattr(graph, "_pg_device", None) or {}).values():
            t.record_stream(main)
        return graph

    def step_resident(self):
        graph = self._build_on_side_stream(lambda: self.d_buf, materialise_host=False)
        out = self.train_and_extract(graph)
        graph.node_sequences                       # the node names are part of the step: decoded here, under the GPU work
        return out + (graph,)

    def _finish_e2e(self):
        """Host-side results of the step whose DirectGCN replay is in flight on the main stream."""
        if self.pending is None:
            return None
        loss, emb, graph = self.pending
        self.pending = None
        graph.node_sequences                                                      # node names decoded under the GPU work
        graph.wait_ready()                                                        # D2H of the five matrices (side stream) has landed
        emb_host = emb.cpu().numpy()                                              # D2H: embeddings (models_utils.py:265-273)
        return float(loss.item()), emb_host, graph

    def step_e2e(self):
        if self.h_buf is None:
            # host format of the corpus: 5 bits per symbol (8 symbols in 5 bytes; what the ingest keeps in host memory for corpora
            # that are streamed to the GPU), unpacked on the device right after the copy; --host-bytes keeps 1 byte per symbol
            host = self.d_buf.cpu()
            packed = None if self.host_bytes else self.corpus.pack5(host)
            self.h_packed = packed
            self.h_buf = packed.packed if packed is not None else host.pin_memory()
            self.d_unpacked = torch.empty(self.nbytes, dtype=torch.uint8, device=self.dev) if packed is not None else None
            self.up = self.corpus.CorpusUploader(self.dev)
            self.up.submit(self.h_buf)

        def prefetch_next():                                                     # H2D of the NEXT step's corpus bytes, issued every
            self.up.release()                                                    # step: it runs on the copy stream under this
            self.up.submit(self.h_buf)                                           # step's extract/normalise/DirectGCN kernels

        def acquire():
            d = self.up.acquire()                                                # the building stream waits for the copy
            return d if self.h_packed is None else self.corpus.unpack5(d, self.h_packed.n_symbols, self.d_unpacked)

        # H2D of this step's bytes; graph object on the host (reference contract)
        graph = self._build_on_side_stream(acquire, materialise_host=True, after_count=prefetch_next)
        if not self.pipelined:
            self.pending = self.train_and_extract(graph) + (graph,)
            return self._finish_e2e()
        prev = self._finish_e2e()        # batch k-1: its replay ran on the main stream while batch k was being built
        self.pending = self.train_and_extract(graph) + (graph,)
        return prev

    def flush(self):
        """Results of the last pipelined e2e step (inside the timed region: every batch's outputs reach the host)."""
        return self._finish_e2e()

    def collect_count_ms(self):
        if getattr(self, "_pending_count_event", False):
            self.ev[1].synchronize()
            self.count_ms.append(self.ev[0].elapsed_time(self.ev[1]))
            self._pending_count_event = False


def next_node_labels(a_out: torch.Tensor, n: int) -> torch.Tensor:
    """argmax-weight successor per node (reference protgram_directgcn_trainer.py:222-237; ties -> first;
    nodes without successors -> themselves).  Setup only, untimed in both arms."""
    idx, val = a_out.indices(), a_out.values()
    labels = torch.arange(n, dtype=torch.int64)
    order = torch.argsort(val, stable=True)  # ascending weight: the last write per row wins = max
    labels[idx[0][order]] = idx[1][order]
    return labels


def rmat_edges(pipe, log2_nodes, edges_per_node, keep_rows=None, chunk_log2=26):
    """R-MAT power-law digraph (a,b,c,d = .57,.19,.19,.05; integer weights 1..7), seeded identically on every
    rank; raw edge list (duplicates included) with randomly relabelled node ids.  Generated in chunks of 2^chunk_log2 edges
    (one counter-seeded generator per chunk: the list does not depend on the chunking of other ranks); keep_rows = (lo, hi)
    keeps only the edges that START in those rows, so a rank of config C5's 1.07 G-edge graph never holds the whole list."""
    dev = pipe.dev
    n = 1 << log2_nodes
    e = n * edges_per_node
    # R-MAT puts every hub at the low ids; relabel the nodes with a seeded random permutation (what any 1-D
    # partitioner does for power-law graphs) so that equal row blocks carry equal numbers of nonzeros
    perm = torch.randperm(n, generator=torch.Generator(device=dev).manual_seed(SEED), device=dev)
    parts = []
    step = min(e, 1 << chunk_log2)
    for c, first in enumerate(range(0, e, step)):
        m = min(step, e - first)
        g = torch.Generator(device=dev).manual_seed(SEED * 1_000_003 + c + 1)
        src = torch.zeros(m, dtype=torch.int64, device=dev)
        dst = torch.zeros(m, dtype=torch.int64, device=dev)
        for _ in range(log2_nodes):  # quadrant -> (src bit, dst bit) = 00,01,10,11
            r = torch.rand(m, generator=g, device=dev)
            src.mul_(2).add_((r >= 0.76).to(torch.int64))
            dst.mul_(2).add_((((r >= 0.57) & (r < 0.76)) | (r >= 0.95)).to(torch.int64))
            del r
        w = torch.randint(1, 8, (m,), generator=g, device=dev).to(torch.float32)
        src, dst = perm[src], perm[dst]
        if keep_rows is not None:
            sel = (src >= keep_rows[0]) & (src < keep_rows[1])
            src, dst, w = src[sel], dst[sel], w[sel]
            del sel
        parts.append((src, dst, w))
    del perm
    src, dst, w = (torch.cat([p[k] for p in parts]) for k in range(3)) if len(parts) > 1 else parts[0]
    return n, e, src, dst, w


def rmat_graph(pipe, log2_nodes, edges_per_node):
    """The R-MAT graph pushed through the reference normalisation (coalesce + a8/a9) on the device -> shared-pattern CSR."""
    gu = pipe.gu
    n, e, src, dst, w = rmat_edges(pipe, log2_nodes, edges_per_node)
    s, d, wv = gu.device_coalesce(src, dst, w, n)
    del src, dst, w
    res = gu.device_normalize(s, d, wv, n, 1e-9)
    del s, d, wv
    for k in ("in_src", "in_dst", "in_w"):
        res.pop(k)
    torch.cuda.empty_cache()
    return n, e, res


def rmat_row_block(pipe, dist, log2_nodes, edges_per_node):
    """This rank's row block of the same graph WITHOUT ever normalising the whole graph on one GPU:
    the rank keeps the R-MAT edges that start in its rows, coalesces them, and `normalize_row_partitioned` does the rest
    (all-to-all of edges by owner of the target, all-gathered degree vectors, row-block kernels).
    -> (n, e, result dict, ms of the partitioned normalisation as max over ranks)."""
    from protgram_directgcn_b200.host import partitioned as part
    gu, dev = pipe.gu, pipe.dev
    lo, hi, per = part.row_range(1 << log2_nodes, pipe.rank, pipe.world)
    n, e, src, dst, w = rmat_edges(pipe, log2_nodes, edges_per_node, keep_rows=(lo, hi))
    s, d, wv = gu.device_coalesce(src, dst, w, n)
    del src, dst, w
    torch.cuda.empty_cache()
    best, res = None, None
    for _ in range(2):  # the first call also warms NCCL's all-to-all channels
        del res
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        res = part.normalize_row_partitioned(s, d, wv, n, 1e-9, dist.group.WORLD)
        b.record()
        torch.cuda.synchronize()
        best = a.elapsed_time(b)
    t = torch.tensor([best], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res["unique_out_edges_local"] = int(s.numel())
    res["recv_in_edges_local"] = int(res["in_w"].numel())
    for k in ("in_src", "in_dst", "in_w", "rs_out", "rs_in", "deg"):
        res.pop(k)
    del s, d, wv
    torch.cuda.empty_cache()
    return n, e, res, float(t.item())


def _time_ms(fn, iters, warm=2):
    for _ in range(warm):
        fn()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return statistics.mean(a.elapsed_time(b) for a, b in evs)


def measured_traffic(kernel_prefix, fname="r02_spmm_traffic.json"):
    """DRAM bytes per call (all launches of the call) of an SpMM kernel family from the committed ncu capture
    (profiles/r02_spmm_traffic.json, written by tools/ncu_traffic.py from `ncu --set full` of tools/run_kernels.py spmm:
    the same graph and shapes as this leg).  None if the capture is missing."""
    path = os.path.join(ROOT, "profiles", fname)
    if not os.path.exists(path):
        return None
    with open(path) as f:
        rec = json.load(f)
    seen = {}
    for l in rec["launches"]:
        if l["kernel"].startswith(kernel_prefix):
            seen.setdefault(l["grid"], l["dram_read_bytes"] + l["dram_write_bytes"])   # one rows-mode + one items-mode launch per call
    return {"bytes_per_call": sum(seen.values()), "launches_per_call": len(seen), "source": f"profiles/{fname} <- " + rec["source"]} if seen else None


def spmm_large_leg(pipe, peak_gbs, log2_nodes=21, edges_per_node=16, F=128, iters=5):
    """DirectGCN propagation where X does not fit L2: R-MAT power-law digraph, hidden 128 (config C5's
    per-GPU shape scaled to one GPU).  fanout_fwd = the forward SpMM; fanout_scaled_bwd = the propagation of the layer's
    backward (gate-scaled fan-out of dY); fanin_bwd_operator = gradient of the bare operator (gathers the 3F-wide dZ)."""
    nat, dev = pipe.nat, pipe.dev
    n, e, res = rmat_graph(pipe, log2_nodes, edges_per_node)
    P = int(res["pattern_nnz"])
    x = torch.randn(n, F, device=dev)
    z = torch.empty(n, 3 * F, device=dev)
    y = torch.empty(n, F, device=dev)
    gates = [torch.rand(n, device=dev) + 0.5 for _ in range(3)]
    st = nat.stream_ptr()
    args = (nat.ptr(res["rowptr"]), nat.ptr(res["col"]), nat.ptr(res["val_in"]), nat.ptr(res["val_out"]), nat.ptr(res["val_und"]), 3, n, F)
    plan = nat.SpmmPlan(res["rowptr"])
    fo = lambda: nat.call("pg_spmm_fanout", *args, nat.ptr(x), F, nat.ptr(z), 3 * F, 0, plan.ref(3 * F), st)
    fs = lambda: nat.call("pg_spmm_fanout_scaled", *args, nat.ptr(x), F, nat.ptr(z), 3 * F, 0, nat.ptr(gates[0]), nat.ptr(gates[1]),
                          nat.ptr(gates[2]), 1, plan.ref(3 * F), st)
    fi = lambda: nat.call("pg_spmm_fanin", *args, nat.ptr(z), 3 * F, 0, None, 0, nat.ptr(y), F, 0, plan.ref(3 * F), st)
    out = {"graph": f"R-MAT 2^{log2_nodes} nodes (ids randomly relabelled), {e} directed edges before dedupe, pattern nnz {P}", "F": F, "pattern_nnz": P,
           "nodes": n}
    deg = (res["rowptr"][1:] - res["rowptr"][:-1])
    out["max_row_nnz"] = int(deg.max())
    out["long_rows"] = {"chunk": plan.chunk, "rows": plan.n_long, "slices": plan.n_items}
    # algorithmic bytes, SURVEY 8(d) with G = P (X = n * F * 4 bytes is far beyond L2): rowptr + col + 3 values + one F-wide row
    # fetch per stored entry + the output rows
    fo_bytes = 8 * (n + 1) + 16 * P + 4 * F * P + 12 * n * F
    for name, fn, bytes_alg, fam in (("fanout_fwd", fo, fo_bytes, "spmm_fanout"), ("fanout_scaled_bwd", fs, fo_bytes + 12 * P, "spmm_fanout"),
                                     ("fanin_bwd_operator", fi, 8 * (n + 1) + 16 * P + 12 * F * P + 4 * n * F, "spmm_fanin")):
        ms = _time_ms(fn, iters)
        out[name] = {"ms": ms, "edges_per_s": 3 * P / (ms * 1e-3), "algorithmic_bytes": bytes_alg,
                     "achieved_gbs": bytes_alg / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": bytes_alg / (ms * 1e-3) / 1e9 / peak_gbs}
        tr = measured_traffic(fam) if (log2_nodes, edges_per_node, F) == (21, 16, 128) and name != "fanout_scaled_bwd" else None
        if tr:
            out[name].update({"dram_traffic_bytes_ncu": tr["bytes_per_call"], "dram_gbs_by_ncu_traffic": tr["bytes_per_call"] / (ms * 1e-3) / 1e9,
                              "frac_of_hbm_peak_by_ncu_traffic": tr["bytes_per_call"] / (ms * 1e-3) / 1e9 / peak_gbs, "traffic_source": tr["source"]})
    return out


def spmm_partitioned_leg(pipe, peak_gbs, dist, log2_nodes_per_gpu=21, edges_per_node=16, F=128, iters=5, compare_allgather=True,
                         parity_rows=256, transport=None):
    """Config C5's shape: R-MAT graph with 2^log2_nodes_per_gpu nodes PER GPU, hidden 128, rows partitioned over the ranks
    (SURVEY 8e).  Every rank generates the seeded edge list, keeps the edges that start in its rows and gets its block of the
    propagation matrices from the row-partitioned normalisation (`normalise_partitioned`, timed, max over ranks).
      fwd  = halo exchange of X (rows this block references: pack kernel + all_to_all_single over NVLink, pipelined over
             feature-column chunks) + local fan-out SpMM over [own rows | halo rows]
      bwd  = the propagation of the LAYER's backward: the same exchange on the F-wide dY plus the three gates of every halo
             row, local gate-scaled fan-out (dX = sum_v (A_v diag(g_v) dY) W_v^T; the 3F-wide gated gradient never moves)
      bwd_operator = gradient of the bare operator Z = fan-out(X) given dZ [3F wide]: halo exchange of dZ + local fan-in
    Times are max over ranks; edges/s is the whole job (3 x pattern nnz of the full graph per pass).  `allgather_r1` is the
    round-1 exchange (all_gather_into_tensor of every row) for comparison.  `parity` holds sampled rows of this rank's block
    against an fp64 evaluation of the same CSR rows over the all-gathered operand."""
    import math
    from protgram_directgcn_b200.host import partitioned as part
    nat, dev, world = pipe.nat, pipe.dev, pipe.world
    log2_nodes = log2_nodes_per_gpu + int(round(math.log2(world)))
    n, e, res, ms_norm = rmat_row_block(pipe, dist, log2_nodes, edges_per_node)
    group = dist.group.WORLD
    saved_transport = part.HALO_TRANSPORT
    if transport is not None:
        part.HALO_TRANSPORT = transport
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    prop = part.RowPartitionedPropagation.from_local(part.local_csr(res), n, group=group, symmetric=True)
    torch.cuda.synchronize()
    plan_s = time.perf_counter() - t0
    halo, csr_ext = prop.halo, prop.local_ext
    p_local = int(prop.local.col.numel())
    tot = torch.tensor([p_local, res["unique_out_edges_local"], halo.num_halo, halo.num_serve], device=dev, dtype=torch.int64)
    dist.all_reduce(tot)
    P, unique_edges = int(tot[0]), int(tot[1])
    norm_info = {"ms": ms_norm, "unique_edges": unique_edges, "edges_per_s": unique_edges / (ms_norm * 1e-3),
                 "exchange": "all_to_all_single of (src, dst, w) by owner of the target (20 B per edge) + all_gather of the "
                             "weighted out-/in-degrees (fp64) and the undirected degree (int32)",
                 "note": "each rank starts from the coalesced out-edges of its own rows; the whole graph never exists on one GPU"}
    del res
    torch.cuda.empty_cache()
    per = prop.per
    g = torch.Generator(device=dev).manual_seed(SEED + pipe.rank)
    x_local = torch.randn(per, F, device=dev, generator=g)
    dy_local = torch.randn(per, F, device=dev, generator=g)
    dz_local = torch.randn(per, 3 * F, device=dev, generator=g)
    gates = tuple(torch.rand(per, device=dev, generator=g) + 0.5 for _ in range(3))
    mx = lambda v: (lambda t: (dist.all_reduce(t, op=dist.ReduceOp.MAX), float(t.item()))[1])(torch.tensor([v], device=dev, dtype=torch.float64))
    holder = {}

    def fwd():
        holder["z"] = part._halo_fanout(csr_ext, halo, x_local, F)

    def bwd():
        holder["t"] = part._halo_fanout(csr_ext, halo, dy_local, F, scales=gates, scale_stride=1)

    def bwd_op():
        holder["dx"] = part._halo_fanin(csr_ext, halo, dz_local, F, None)

    def comm():
        holder["h"] = halo.exchange(x_local)

    def local_fo():
        nat.call("pg_spmm_fanout_split", nat.ptr(csr_ext.rowptr), nat.ptr(csr_ext.col), nat.ptr(csr_ext.vals[0]), nat.ptr(csr_ext.vals[1]),
                 nat.ptr(csr_ext.vals[2]), 3, per, F, nat.spmm_operand(x_local, holder["h"], per), nat.ptr(holder["z"]), 3 * F, 0, F,
                 None, None, None, 0, csr_ext.plan(3 * F), nat.stream_ptr())

    recv_rows = halo.num_halo
    out = {"graph": f"R-MAT 2^{log2_nodes} nodes ({world} x 2^{log2_nodes_per_gpu}, ids randomly relabelled), {e} directed edges before dedupe, pattern nnz {P}",
           "nodes": n, "F": F, "pattern_nnz": P, "rows_per_gpu": per, "partition": "1-D rows; block columns renumbered into [own rows | halo rows]",
           "local_nnz_max_over_mean": mx(float(p_local)) / (P / world),
           "exchange": ("halo, peer-memory transport: ONE kernel (pg_halo_push) stores the referenced rows into the peers' CUDA-IPC mapped receive "
                        "slots over NVLink and publishes an epoch flag; pg_halo_wait in front of the SpMM" if halo._peer_transport() is not None else
                        "halo, NCCL transport: pg_gather_rows + all_to_all_single of the referenced rows") +
                       f", then the local SpMM over [own rows | halo rows] ({len(part._feature_chunks(F, per, world))} feature-column chunk(s))",
           "halo": {"rows_received_per_gpu_mean": int(tot[2]) / world, "rows_served_per_gpu_mean": int(tot[3]) / world,
                    "fraction_of_remote_rows": (int(tot[2]) / world) / max(1, n - per), "plan_build_s_once_per_graph": mx(plan_s)},
           "normalise_partitioned": norm_info}
    for name, fn, width, extra_cols in (("fwd", fwd, F, 0), ("bwd", bwd, F, 4), ("bwd_operator", bwd_op, 3 * F, 0)):
        dist.barrier()
        ms = mx(_time_ms(fn, iters))
        recv = 4 * (width + extra_cols) * recv_rows
        out[name] = {"ms": ms, "edges_per_s": 3 * P / (ms * 1e-3), "nvlink_recv_bytes_this_gpu": recv,
                     "all_gather_would_receive_bytes": 4 * width * per * (world - 1)}
    # the feature-column pipelining of the exchange (off by default, see host/partitioned.py): measured for the record
    saved = (part.PIPELINE_CHUNKS, part.PIPELINE_MIN_BYTES)
    try:
        part.PIPELINE_CHUNKS, part.PIPELINE_MIN_BYTES = 2, 0
        dist.barrier()
        out["fwd"]["ms_with_2_feature_chunks_pipelined"] = mx(_time_ms(fwd, iters))
    finally:
        part.PIPELINE_CHUNKS, part.PIPELINE_MIN_BYTES = saved
    dist.barrier()
    ms_comm = mx(_time_ms(comm, iters))
    fwd()
    comm()
    dist.barrier()
    ms_local = mx(_time_ms(local_fo, iters))
    recv = 4 * F * recv_rows
    alg_local = 8 * (per + 1) + 16 * p_local + 4 * F * p_local + 12 * per * F
    out["fwd"].update({"ms_exchange_alone": ms_comm, "ms_local_spmm_alone": ms_local, "overlap_hidden_ms": ms_comm + ms_local - out["fwd"]["ms"],
                       "nvlink_gbs_this_gpu_exchange_alone": recv / (ms_comm * 1e-3) / 1e9, "nvlink_frac_of_measured_770": recv / (ms_comm * 1e-3) / 1e9 / 770.0,
                       "local_spmm_algorithmic_bytes": alg_local, "local_spmm_gbs": alg_local / (ms_local * 1e-3) / 1e9,
                       "local_spmm_frac_of_hbm_peak": alg_local / (ms_local * 1e-3) / 1e9 / peak_gbs})
    # one whole DirectGCN layer (F -> F, identity residual, tensor-core transform) on the block, forward + backward: in the backward
    # the halo exchange of dY is posted before the weight- and gate-gradient GEMMs and waited for after them
    try:
        import protgram_directgcn_b200 as pg
        data = part.partitioned_data(x_local, prop.local, n, group, halo=prop.halo)
        st_obj = data.edge_index_in._pg_struct
        layer = pg.DirectGCNLayer(F, F, per, True).to(dev)
        edges = (data.edge_index_in, data.edge_weight_in, data.edge_index_out, data.edge_weight_out, data.edge_index_undirected_norm,
                 data.edge_weight_undirected_norm)
        xr = x_local.clone().requires_grad_(True)
        dh = torch.randn(per, F, device=dev, generator=g)

        def layer_step():
            xr.grad = None
            layer.zero_grad(set_to_none=True)
            layer._run(xr, edges, None, None, None, True, 0.01).backward(dh)

        dist.barrier()
        ms_layer = mx(_time_ms(layer_step, iters))
        begin = type(st_obj).fanout_begin
        try:
            type(st_obj).fanout_begin = lambda self, *a, **k: None          # same step with the exchange NOT overlapped
            dist.barrier()
            ms_layer_serial = mx(_time_ms(layer_step, iters))
        finally:
            type(st_obj).fanout_begin = begin
        out["layer_fwd_bwd"] = {"what": f"DirectGCNLayer({F}, {F}) forward + backward on the row block (fan-out SpMM, tcgen05 transform, weight / gate / "
                                        "input gradients), exchanges included", "ms": ms_layer, "ms_backward_exchange_not_overlapped": ms_layer_serial,
                                "edges_per_s": 2 * 3 * P / (ms_layer * 1e-3)}
        del data, layer, edges, xr, dh, st_obj
    except Exception as exc:  # noqa: BLE001
        out["layer_fwd_bwd"] = {"error": repr(exc)}
    torch.cuda.empty_cache()
    x_full = None
    if compare_allgather and world * per * F * 4 <= 40 << 30:
        def ag_fwd():
            holder["xf"] = part._all_gather_rows(x_local, group)
            holder["z1"] = part._spmm_fanout(prop.local, holder["xf"], per, F)
        dist.barrier()
        ms_ag = mx(_time_ms(ag_fwd, iters))
        out["allgather_r1"] = {"fwd_ms": ms_ag, "speedup_of_halo_exchange": ms_ag / out["fwd"]["ms"],
                               "bitwise_equal_to_halo_path": bool(torch.equal(holder["z1"], holder["z"]))}
        x_full = holder.pop("xf")
        holder.pop("z1")
    if parity_rows and world * per * F * 4 <= 40 << 30:
        if x_full is None:
            x_full = part._all_gather_rows(x_local, group)
        rows = torch.randperm(min(per, max(1, prop.hi - prop.lo)), device=dev, generator=g)[:parity_rows]
        rp = prop.local.rowptr
        worst = 0.0
        for r in rows.tolist():
            b, e_ = int(rp[r]), int(rp[r + 1])
            cols = prop.local.col[b:e_].long()
            xr = x_full[cols].double()
            for v in range(3):
                ref = (prop.local.vals[v][b:e_].double().unsqueeze(1) * xr).sum(0)
                got = holder["z"][r, v * F:(v + 1) * F].double()
                worst = max(worst, float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30)))
        out["parity"] = {"rows_sampled_per_rank": int(rows.numel()), "max_rel_err_vs_fp64_over_ranks": mx(worst), "bar": 1e-4}
        del x_full
    holder.clear()
    tr = halo._peer_transport()
    if tr is not None:
        tr.check()
    halo.close()                      # collective: the peer-memory receive rings are cudaMalloc'ed outside torch's allocator
    part.HALO_TRANSPORT = saved_transport
    torch.cuda.empty_cache()
    return out


def write_bench_fasta(pipe, path):
    """The bench corpus of this rank as a FASTA file (UniProt-style headers `>sp|P0000001|SYN`, one sequence line per record)."""
    host = pipe.d_buf.cpu().numpy()
    lead = 1 if pipe.rank == 0 else 0
    seqs = host[lead:lead + NSEQ * (SEQ_LEN + 2)].reshape(NSEQ, SEQ_LEN + 2)[:, :SEQ_LEN]
    ids = np.char.zfill(np.arange(NSEQ).astype("U7"), 7)
    head = np.frombuffer("".join(f">sp|P{i}|SYN\n" for i in ids).encode("ascii"), dtype=np.uint8).reshape(NSEQ, 17)
    rec = np.concatenate([head, seqs, np.full((NSEQ, 1), 10, dtype=np.uint8)], axis=1)
    rec.tofile(path)
    return int(rec.size)


def e2e_api_leg(pipe, steps=3):
    """The calls a user of the reference makes, end to end FROM A FASTA FILE (VERDICT r1 #3b):
        GraphBuilder(config).run()           parse + pack the file, H2D, n = 1..3 graphs, five matrices each to the host,
                                             three pickles (+ CSR sidecars) written            (run_graph_builder.py:36-56)
        DataUtils.load_object(n = 3 pickle)  -> graph.gcn_data(x, device)                      (trainer :288-299, :362-367)
        labels, ProtGramDirectGCN(...), Adam; one epoch of the reference's full-batch loop: model(data) -> F.nll_loss + L2 term
        -> backward -> step; extract_gcn_node_embeddings -> numpy                                 (trainer :76-108, :380)
    Wall clock per step (device synchronised at the end), nothing prepared outside the step except the file itself
    (written to tmpfs / page cache).  This is the slow, file-and-pickle-bound way through the same kernels; the pipelined
    `e2e` figure keeps the corpus bytes and the graph in memory."""
    import contextlib
    import io
    import shutil
    pg = pipe.pg
    base = tempfile.mkdtemp(prefix="pgb200_api_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    out = {}
    try:
        fasta = os.path.join(base, "corpus.fasta")
        nbytes = write_bench_fasta(pipe, fasta)
        cfg = pg.Config()
        cfg.BASE_OUTPUT_DIR = os.path.join(base, "out")
        cfg.GCN_INPUT_FASTA_PATH = fasta
        cfg.GRAPH_OBJECTS_DIR = os.path.join(base, "out", "graphs")
        cfg.GCN_NGRAM_MAX_N, cfg.GRAPH_BUILDER_WORKERS = N_LEVEL, 1
        phases = {"graph_builder_run": [], "load_pickle_and_gcn_data": [], "labels_model_optimizer": [], "train_epoch": [], "extract_embeddings": []}
        total = []
        sink = io.StringIO()
        for it in range(steps + 1):
            shutil.rmtree(cfg.GRAPH_OBJECTS_DIR, ignore_errors=True)
            torch.cuda.synchronize()
            t = [time.perf_counter()]
            with contextlib.redirect_stdout(sink):
                pg.GraphBuilder(cfg).run()
            torch.cuda.synchronize(); t.append(time.perf_counter())
            graph = pg.DataUtils.load_object(os.path.join(cfg.GRAPH_OBJECTS_DIR, f"ngram_graph_n{N_LEVEL}.pkl"))
            n = graph.number_of_nodes
            x = torch.randn(n, DIMS[0], generator=torch.Generator().manual_seed(SEED))
            data = graph.gcn_data(x, pipe.dev)
            torch.cuda.synchronize(); t.append(time.perf_counter())
            labels = pg.generate_next_node_labels(graph)[0].to(pipe.dev)
            torch.manual_seed(SEED)
            model = pg.ProtGramDirectGCN(DIMS, n, n, N_LEVEL, 0, 512, DROPOUT, True).to(pipe.dev)
            opt = torch.optim.Adam(model.parameters(), lr=LR)
            torch.cuda.synchronize(); t.append(time.perf_counter())
            model.train()
            opt.zero_grad()
            logp, _ = model(data=data)
            loss = torch.nn.functional.nll_loss(logp, labels) + L2_LAMBDA * sum(p.norm(2).pow(2) for p in model.parameters() if p.requires_grad)
            loss.backward()
            opt.step()
            torch.cuda.synchronize(); t.append(time.perf_counter())
            emb = pg.EmbeddingProcessor.extract_gcn_node_embeddings(model, data, pipe.dev) if hasattr(pg.EmbeddingProcessor, "extract_gcn_node_embeddings") else None
            if emb is None:
                model.eval()
                with torch.no_grad():
                    emb = model(data=data)[1].cpu().numpy()
            torch.cuda.synchronize(); t.append(time.perf_counter())
            if it > 0:      # first pass = warm-up (allocator, page cache, CUDA context of the loss kernels)
                for k, a, b in zip(phases, t[:-1], t[1:]):
                    phases[k].append((b - a) * 1e3)
                total.append((t[-1] - t[0]) * 1e3)
            del graph, data, model, opt
        ms = statistics.mean(total)
        out = {"what": "GraphBuilder(config).run() on a FASTA file (n = 1..3, pickles) + pickle load + one epoch of the reference's full-batch "
                       "training loop + embedding extraction, through the package's public classes only",
               "fasta_bytes": nbytes, "steps": steps, "ms_per_step": ms, "residues_per_s": NSEQ * SEQ_LEN / (ms * 1e-3),
               "phases_ms": {k: round(statistics.mean(v), 2) for k, v in phases.items()}, "timing": "wall clock, device synchronised after every phase",
               "nodes": n, "embedding_shape": list(np.asarray(emb).shape), "loss": float(loss)}
    finally:
        shutil.rmtree(base, ignore_errors=True)
    return out


def pooling_leg(pipe, emb, graph, iters=5, cpu_sample=400):
    """Row f2: protein-level pooling of the level's embeddings over the whole corpus (reference
    models_utils.py:210-262: a Python loop over every residue).  Device-resident raw sequences (the bench corpus
    without its padding bytes), n = N_LEVEL, F = DIMS[-1]; the oracle is timed on `cpu_sample` proteins."""
    from oracle import next_oracle, ngram_oracle
    nat, dev = pipe.nat, pipe.dev
    lead = 1 if pipe.rank == 0 else 0
    raw = pipe.d_buf[lead:lead + NSEQ * (SEQ_LEN + 2)].view(NSEQ, SEQ_LEN + 2)[:, :SEQ_LEN].contiguous().view(-1)
    offsets = torch.arange(NSEQ + 1, dtype=torch.int64, device=dev) * SEQ_LEN
    nodes = graph.node_sequences
    symbols = np.array(sorted({ord(c) for s in nodes for c in s}), dtype=np.uint8)
    rank = np.full(256, 255, dtype=np.uint8)
    rank[symbols] = np.arange(symbols.size, dtype=np.uint8)
    codes = np.zeros(len(nodes), dtype=np.int64)
    chars = np.frombuffer("".join(nodes).encode("ascii"), dtype=np.uint8).reshape(len(nodes), N_LEVEL)
    for k in range(N_LEVEL):
        codes = codes * symbols.size + rank[chars[:, k]]
    table = np.full(int(symbols.size) ** N_LEVEL, -1, dtype=np.int32)
    table[codes] = np.arange(len(nodes), dtype=np.int32)
    d_rank, d_tab = torch.from_numpy(rank).to(dev), torch.from_numpy(table).to(dev)
    emb = emb.detach().contiguous()
    F = int(emb.shape[1])
    out = torch.empty((NSEQ, F), dtype=torch.float32, device=dev)
    valid = torch.empty(NSEQ, dtype=torch.uint8, device=dev)
    fn = lambda: nat.call("pg_pool_proteins", nat.ptr(raw), nat.ptr(offsets), NSEQ, N_LEVEL, nat.ptr(d_rank), int(symbols.size),
                          nat.ptr(d_tab), nat.ptr(emb), emb.stride(0), F, nat.ptr(out), out.stride(0), nat.ptr(valid), nat.stream_ptr())
    ms = _time_ms(fn, iters)
    seqs = [(str(i), s) for i, s in enumerate(ngram_oracle.synth_sequences(pipe.rank * NSEQ, cpu_sample, SEQ_LEN, SEED))]
    emb_h = emb.cpu().numpy()
    t0 = time.perf_counter()
    _, pooled, ok = next_oracle.pool_proteins(seqs, N_LEVEL, {s: i for i, s in enumerate(nodes)}, emb_h)
    t_cpu = time.perf_counter() - t0
    exact = bool(np.array_equal(out[:cpu_sample].cpu().numpy()[ok], pooled[ok]))
    return {"what": f"pool_ngram_embeddings_for_protein_fast over {NSEQ} proteins x {SEQ_LEN} residues, n={N_LEVEL}, F={F}",
            "ms": ms, "residues_per_s": NSEQ * SEQ_LEN / (ms * 1e-3), "proteins_per_s": NSEQ / (ms * 1e-3),
            "cpu_oracle": {"proteins": cpu_sample, "seconds": t_cpu, "residues_per_s": cpu_sample * SEQ_LEN / t_cpu, "cores": 1},
            "bit_exact_vs_oracle_on_sample": exact}


def layer_gemm_leg(pipe, n=160_000, f_in=256, f_out=256, iters=10):
    """The dense transform of one DirectGCN layer at config C3's shape (N = 160k, hidden 256, K_ext = 771): the
    tcgen05 kernel (3 x TF32 split, fp32-level accuracy) against the SIMT fp32 kernel.  flops = 2 N K_ext F_out;
    the tensor-pipe figure counts the 3 MMAs per product (what the hardware executes)."""
    nat, dev = pipe.nat, pipe.dev
    k_ext = 3 * f_in + 3
    z, x = torch.randn(n, 3 * f_in, device=dev), torch.randn(n, f_in, device=dev)
    g = [torch.rand(n, device=dev) + 0.5 for _ in range(3)]
    w = torch.randn(k_ext, f_out, device=dev) * 0.1
    c = torch.randn(n, f_out, device=dev)
    h = torch.empty(n, f_out, device=dev)
    ws = torch.empty(nat.query("pg_layer_gemm_fwd_tc_ws_bytes", f_in, f_out, 0), dtype=torch.uint8, device=dev)
    st = nat.stream_ptr()
    simt = lambda: nat.call("pg_layer_gemm_fwd", nat.ptr(z), 3 * f_in, nat.ptr(x), f_in, nat.ptr(g[0]), nat.ptr(g[1]), nat.ptr(g[2]), 1,
                            nat.ptr(w), nat.ptr(c), f_out, n, f_in, f_out, 0, 1, 0.01, nat.ptr(h), f_out, st)
    tc = lambda: nat.call("pg_layer_gemm_fwd_tc", nat.ptr(z), 3 * f_in, nat.ptr(x), f_in, nat.ptr(g[0]), nat.ptr(g[1]), nat.ptr(g[2]), 1,
                          nat.ptr(w), nat.ptr(c), f_out, n, f_in, f_out, 0, 1, 0.01, nat.ptr(h), f_out, nat.ptr(ws), ws.numel(), st)
    ms_simt, ms_tc = _time_ms(simt, iters, warm=3), _time_ms(tc, iters, warm=3)
    nat.call("pg_layer_gemm_fwd_tc_check", nat.ptr(ws), f_in, f_out, 0, st)
    flops = 2.0 * n * k_ext * f_out
    return {"shape": f"N={n} F_in={f_in} F_out={f_out} K_ext={k_ext} (C3 layer)", "simt_fp32_ms": ms_simt,
            "simt_fp32_tflops": flops / ms_simt / 1e9, "tcgen05_3xtf32_ms": ms_tc, "tcgen05_effective_tflops": flops / ms_tc / 1e9,
            "tcgen05_tensor_pipe_tflops_tf32": 3 * flops / ms_tc / 1e9,
            "note": "tf32 dense peak is half the bf16 one (MEASURED_PEAKS bf16_tflops / 2); every product costs 3 tf32 MMAs"}


def _max_over_ranks(dist, dev, v):
    if dist is None:
        return float(v)
    t = torch.tensor([v], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _build_key_range_variant(pipe, dist, buf, n_level, d_rank, sigma, ws, iters, nodes_expected, edges_expected, transitions_expected):
    """The same build with the merge the north star words literally: REDUCE-SCATTER of the dense tables over key ranges, then
    every rank extracts only the edges of its own key range (whole source rows; `csrc/extract_range.cu`) -- the merged table and
    the edge list never exist on one GPU.  One small extra collective (the sigma^n-byte presence table, MAX) gives every rank
    the same node numbering.  Timed like the replicated variant (count -> merge -> extract, max over ranks, median of passes)."""
    db, dev, rank, world = pipe.db, pipe.dev, pipe.rank, pipe.world
    pow_n, pow_m = db.table_sizes(n_level, sigma)
    codes_per = (pow_n + world - 1) // world
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t_count, t_merge, t_extract = [], [], []
    res = None
    for it in range(iters + 1):
        res = None
        padded = torch.zeros(world * codes_per * sigma, dtype=torch.int64, device=dev)
        short = torch.zeros(pow_n, dtype=torch.uint8, device=dev)
        dist.barrier()
        torch.cuda.synchronize()
        ev[0].record()
        db.count_level(buf, n_level, d_rank, sigma, padded[:pow_m], short, ws)
        ev[1].record()
        local = db.merge_tables_by_key_range(padded, codes_per, sigma, pipe.group)
        s32 = short.to(torch.int32)
        dist.all_reduce(s32, op=dist.ReduceOp.MAX)
        short = s32.to(torch.uint8)
        ev[2].record()
        res = db.extract_key_range(local, short, n_level, sigma, rank * codes_per, codes_per, pipe.group)
        ev[3].record()
        torch.cuda.synchronize()
        if it > 0:
            t_count.append(ev[0].elapsed_time(ev[1])); t_merge.append(ev[1].elapsed_time(ev[2])); t_extract.append(ev[2].elapsed_time(ev[3]))
        del padded, short, local
    node_code, src, dst, cnt = res
    tot = torch.stack([torch.tensor(int(src.numel()), device=dev, dtype=torch.int64), cnt.sum()])
    dist.all_reduce(tot)
    mc, mm, me = (_max_over_ranks(dist, dev, statistics.median(t)) for t in (t_count, t_merge, t_extract))
    return {"merge": "reduce_scatter_tensor over key ranges (source codes split evenly) + all_reduce(MAX) of the presence table",
            "count_ms": mc, "merge_reduce_scatter_ms": mm, "extract_own_key_range_ms": me, "build_ms": mc + mm + me,
            "merge_bytes_per_gpu": world * codes_per * sigma * 8, "presence_bytes": pow_n,
            "edges_on_this_rank": int(src.numel()),
            "matches_replicated_build": bool(int(node_code.numel()) == nodes_expected and int(tot[0]) == edges_expected
                                             and int(tot[1]) == transitions_expected)}


def build_scale_leg(pipe, dist, n_level, total_seqs, iters=3, normalise=True):
    """Graph build of one n level at BASELINE configs C3 (n=4) / C4 (n=5, 50 M sequences = 17.5 G residues):
    `total_seqs` 350-residue sequences split over the ranks by contiguous ranges (STRONG scaling: the corpus is
    fixed), each shard generated on its GPU (counter-based, shard independent); timed = count -> NCCL sum of the
    dense (n+1)-gram tables -> node ids / edge table.  Times are device times, max over ranks."""
    nat, db, dev, rank, world = pipe.nat, pipe.db, pipe.dev, pipe.rank, pipe.world
    per = total_seqs // world
    nbytes = per * (SEQ_LEN + 2) + (1 if rank == 0 else 0)
    buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    nat.call("pg_synth_corpus", nat.ptr(buf), rank * per, per, SEQ_LEN, SEED, int(rank == 0), nat.stream_ptr())
    symbols, d_rank = pipe.corpus.discover_alphabet(buf, pipe.group)
    sigma = int(symbols.size)
    pow_n, pow_m = db.table_sizes(n_level, sigma)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ws = db.count_workspace(n_level, sigma, nbytes, dev)     # allocated once, like GraphBuilder does per level
    t_count, t_merge, t_extract = [], [], []
    res = None
    for it in range(iters + 1):
        res = None                                   # release the previous pass's outputs first: the allocator then reuses their blocks
        bins = torch.zeros(pow_m, dtype=torch.int64, device=dev)
        short = torch.zeros(pow_n, dtype=torch.uint8, device=dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        ev[0].record()
        db.count_level(buf, n_level, d_rank, sigma, bins, short, ws)
        ev[1].record()
        if dist is not None:
            dist.all_reduce(bins, op=dist.ReduceOp.SUM)
            s32 = short.to(torch.int32)
            dist.all_reduce(s32, op=dist.ReduceOp.MAX)
            short = s32.to(torch.uint8)
        ev[2].record()
        res = db.extract_level(bins, short, n_level, sigma)
        ev[3].record()
        torch.cuda.synchronize()
        if it > 0:      # first pass = warm-up
            t_count.append(ev[0].elapsed_time(ev[1])); t_merge.append(ev[1].elapsed_time(ev[2])); t_extract.append(ev[2].elapsed_time(ev[3]))
        del bins, short
    node_code, src, dst, cnt = res
    # median over the passes: the extraction allocates ~3 GB of outputs + workspace per pass and an occasional cudaMalloc
    # (allocator growth, not kernel time) would otherwise dominate a 2 ms phase
    mc, mm, me = (_max_over_ranks(dist, dev, statistics.median(t)) for t in (t_count, t_merge, t_extract))
    residues = per * world * SEQ_LEN
    nodes, edges = int(node_code.numel()), int(src.numel())
    out = {"n": n_level, "sequences": per * world, "residues": residues, "sigma": sigma, "table_bins": pow_m, "nodes": nodes,
           "unique_edges": edges, "transitions_counted": int(cnt.sum().item()), "scaling": "strong (fixed corpus split over ranks)",
           "count_ms": mc, "merge_allreduce_ms": mm, "extract_ms": me, "build_ms": mc + mm + me,
           "build_residues_per_s": residues / ((mc + mm + me) * 1e-3), "count_residues_per_s": residues / (mc * 1e-3),
           "count_variant": ("partition by 2-symbol prefix + shared-memory count per bucket (variant P)"
                             if nat.query("pg_ngram_count_ws_bytes_for", n_level, sigma, nbytes) > nat.query("pg_ngram_count_ws_bytes", n_level, sigma)
                             and nbytes >= 32 << 20 else "L2 RED.ADD.u64 on the dense table"),
           "count_frac_of_1B_per_residue_hbm_bound": nbytes / (mc * 1e-3) / 1e9 / peaks()[0]}
    if dist is not None:
        out["merge_bytes_per_gpu"] = pow_m * 8
        try:
            kr = _build_key_range_variant(pipe, dist, buf, n_level, d_rank, sigma, ws, iters, nodes, edges, out["transitions_counted"])
            kr["build_residues_per_s"] = residues / (kr["build_ms"] * 1e-3)
        except Exception as exc:  # noqa: BLE001 - the replicated numbers above stand on their own
            kr = {"error": repr(exc)}
        out["key_range_variant"] = kr
        if "error" not in kr and db._use_key_range_merge(n_level, sigma, pipe.group):
            # the product's merge policy for this table size (data_builder._use_key_range_merge): the headline numbers of the
            # leg are the key-range path's; the all-reduce path stays beside them
            out["allreduce_variant"] = {k: out[k] for k in ("count_ms", "merge_allreduce_ms", "extract_ms", "build_ms", "build_residues_per_s")}
            out.update({"count_ms": kr["count_ms"], "merge_reduce_scatter_ms": kr["merge_reduce_scatter_ms"],
                        "extract_ms": kr["extract_own_key_range_ms"], "build_ms": kr["build_ms"],
                        "build_residues_per_s": kr["build_residues_per_s"], "count_residues_per_s": residues / (kr["count_ms"] * 1e-3),
                        "merge_policy": "table beyond L2: reduce_scatter_tensor over key ranges, every rank extracts its own range "
                                        "(what GraphBuilder.run / build_level_graph do at this size)"})
            out.pop("merge_allreduce_ms", None)
        else:
            out["merge_policy"] = "table fits L2: all_reduce, every rank extracts the whole graph"
    del buf, ws
    if normalise and rank == 0:
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        w = cnt.to(torch.float32)
        resn = pipe.gu.device_normalize(src, dst, w, nodes, 1e-9)
        b.record()
        torch.cuda.synchronize()
        out["normalise_ms"] = a.elapsed_time(b)
        out["pattern_nnz"] = int(resn["pattern_nnz"])
        out["_graph"] = (nodes, resn)
    torch.cuda.empty_cache()
    return out


def c3_directgcn_leg(pipe, nodes, res, dims=(64, 256, 256, 256), classes=21, iters=5):
    """Config C3's DirectGCN: 3 layers of hidden 256 on the n=4 graph (full batch, fused loss, `classes` labels as
    in the closest_aa task).  One step = train forward + nll + backward + Adam + eval-mode embedding extraction.
    edge unit: one stored nonzero of one propagation matrix per layer pass; a step makes 3 passes
    (train fwd, bwd, eval fwd) over 3 matrices x L layers."""
    from protgram_directgcn_b200.host.protgram_directgcn import Data, register_symmetric_structure
    from protgram_directgcn_b200.host.models_utils import EmbeddingProcessor
    dev = pipe.dev
    torch.manual_seed(SEED)
    model = pipe.pg.ProtGramDirectGCN(list(dims), nodes, classes, 4, 0, 512, DROPOUT, True).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=LR, fused=True)
    x = torch.randn(nodes, dims[0], device=dev)
    y = torch.randint(0, classes, (nodes,), device=dev)
    ei = torch.zeros((2, 1), dtype=torch.int64, device=dev)
    vals = (res["val_in"], res["val_out"], res["val_und"])
    register_symmetric_structure(ei, vals, nodes, res["rowptr"], res["col"])
    data = Data(x=x, edge_index_in=ei, edge_weight_in=vals[0], edge_index_out=ei, edge_weight_out=vals[1],
                edge_index_undirected_norm=ei, edge_weight_undirected_norm=vals[2], num_nodes=nodes)
    P, L = int(res["pattern_nnz"]), len(dims) - 1

    def train():
        model.train()
        opt.zero_grad(set_to_none=True)
        loss = model.nll_loss(data, y)
        loss.backward()
        opt.step()
        return loss

    def fwd_only():
        model.eval()
        with torch.no_grad():
            return EmbeddingProcessor.l2_normalize_torch(model.embed(data), eps=model.l2_eps)

    def step():
        train()
        return fwd_only()

    ms_step = _time_ms(step, iters, warm=2)
    ms_fwd = _time_ms(fwd_only, iters, warm=1)
    out = {"dims": list(dims), "nodes": nodes, "pattern_nnz": P, "classes": classes, "ms_per_step_eager": ms_step,
           "ms_eval_forward": ms_fwd, "ms_train_fwd_bwd_adam_eager": ms_step - ms_fwd,
           "edges_per_s_forward": 3 * P * L / (ms_fwd * 1e-3),
           "dense_flops_per_layer_fwd": [2.0 * nodes * (3 * a + 3 + (a + 1 if a != b else 0)) * b for a, b in zip(dims[:-1], dims[1:])]}
    # the same step captured once into a CUDA graph and replayed (what the C2 headline step does; the reference trains up to
    # 500 epochs per level on one fixed graph)
    ms_graph = None
    try:
        from protgram_directgcn_b200.host.graphed_step import GraphedDirectGCNStep
        torch.manual_seed(SEED)
        model2 = pipe.pg.ProtGramDirectGCN(list(dims), nodes, classes, 4, 0, 512, DROPOUT, True).to(dev)
        opt2 = torch.optim.Adam(model2.parameters(), lr=LR, fused=True, capturable=True)
        gs = GraphedDirectGCNStep(model2, opt2, x, y, nodes, P + 1024, l2_lambda=L2_LAMBDA)
        gs.load_structure(res["rowptr"], res["col"], res["val_in"], res["val_out"], res["val_und"])
        ms_graph = _time_ms(gs.replay, iters, warm=2)
        out["ms_per_step_cuda_graph"] = ms_graph
        out["kernels_per_replay"] = gs.kernels_per_replay
        del gs, model2, opt2
    except Exception as exc:  # noqa: BLE001
        out["cuda_graph_error"] = repr(exc)
    best = min(ms_step, ms_graph) if ms_graph else ms_step
    # SURVEY 8(d) B_layer with G = N (X of an n-gram graph stays L2 resident: compulsory reads only), summed over the layers and
    # the three passes of a step (train forward, backward ~ forward's bytes, eval forward)
    b_step = 0
    for a, b in zip(dims[:-1], dims[1:]):
        fg = min(a, 3 * b)
        b_layer = 4 * (nodes + 1) + 16 * P + 4 * fg * nodes + 4 * nodes * a + 8 * nodes * b + 20 * nodes + 4 * (4 * a * b + 6 * b)
        b_step += 3 * b_layer
    out.update({"ms_per_step": best, "edges_per_s_step": 3 * P * L * 3 / (best * 1e-3),
                "roofline": {"bound": "hbm", "unit": "GB/s", "algorithmic_bytes_per_step": b_step, "achieved": b_step / (best * 1e-3) / 1e9,
                             "frac": b_step / (best * 1e-3) / 1e9 / peaks()[0],
                             "note": "SURVEY 8(d) B_layer, G = N, x 3 passes x L layers: the compulsory bytes of the propagation path; the step also runs "
                                     "the dense transforms (tensor-core bound) and the decoder / loss / Adam, which this byte count does not credit"},
                "mode": "dense transform, data gradient and weight gradient on tcgen05 (3 x TF32); eager = autograd over libpgb200 kernels"})
    return out


def phase_breakdown(pipe, reps=5):
    """Untimed diagnostic: wall-clock per phase of the resident step with a device sync after each phase."""
    import collections
    nat, db = pipe.nat, pipe.db
    acc = collections.OrderedDict()

    def lap(name, t):
        torch.cuda.synchronize()
        now = time.perf_counter()
        acc[name] = acc.get(name, 0.0) + (now - t) * 1e3 / reps
        return now

    for _ in range(reps):
        torch.cuda.synchronize()
        t = time.perf_counter()
        symbols, d_rank = pipe.corpus.discover_alphabet(pipe.d_buf, None)
        t = lap("alphabet", t)
        sigma = int(symbols.size)
        bins, short = db.count_level(pipe.d_buf, N_LEVEL, d_rank, sigma)
        t = lap("count", t)
        node_code, src, dst, cnt = db.extract_level(bins, short, N_LEVEL, sigma)
        t = lap("extract", t)
        names = pipe.corpus.LazyNodeNames(node_code, symbols, N_LEVEL).resolve()
        t = lap("node_names", t)
        graph = pipe.gu.DirectedNgramGraph.from_edge_arrays(names, src, dst, cnt.to(torch.float32), n_value=N_LEVEL,
                                                            assume_coalesced=True, result_device=pipe.dev)
        t = lap("graph_object+normalise", t)
        pipe.train_and_extract(graph)
        t = lap("directgcn_step", t)
    return {k: round(v, 3) for k, v in acc.items()}


def run_b200(args):
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the single JSON line (NCCL prints its banner there)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # one process per GPU: run it (and first-touch its pinned upload buffers) on the CPUs / NUMA node next to that GPU --
    # with 8 ranks the end-to-end step is bound by host memory and PCIe traffic, not by the GPUs
    affinity = None
    try:   # also at N = 1: on a two-socket 8-GPU node an unpinned process may run (and first-touch its pinned buffers) on the far socket
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
        affinity = f"nvml ideal cpus ({len(os.sched_getaffinity(0))} cores)"
        if world > 1:
            torch.set_num_threads(max(1, min(8, len(os.sched_getaffinity(0)) // 8)))    # the ranks of one socket share its cores
    except Exception as exc:  # noqa: BLE001 - best effort
        affinity = f"unchanged ({type(exc).__name__})"
    dist = None
    if world > 1:
        import torch.distributed as dist
        import datetime
        # a rank that fails alone must not leave the others in a collective for NCCL's default 10 minutes
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=int(os.environ.get("PGB200_NCCL_TIMEOUT_S", "150"))))
    pipe = B200Pipeline(rank, world, dev)
    pipe.use_cuda_graph = not args.no_cuda_graph
    pipe.pipelined = not args.no_pipeline
    pipe.host_bytes = not args.host_packed5
    nat = pipe.nat
    peak_gbs, peak_src = peaks()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        last = None
        for _ in range(warmup):
            last = fn()         # keep the previous result alive exactly like the timed loop does (same allocator pattern)
            pipe.collect_count_ms()
        # a cudaMalloc inside the timed region is a device-wide stall (seen as one 4-5 ms step): keep warming up, untimed, until the
        # caching allocator (two streams = two pools) has stopped growing for three steps in a row
        quiet, extra = 0, 0
        while quiet < 3 and extra < 16:
            before = torch.cuda.memory_stats(dev)["num_device_alloc"]
            last = fn()
            pipe.collect_count_ms()
            quiet = quiet + 1 if torch.cuda.memory_stats(dev)["num_device_alloc"] == before else 0
            extra += 1
        pipe.flush()
        pipe.count_ms.clear()
        gc.collect()
        gc.disable()            # timing hygiene (as timeit does): a gen-2 collection is a 50 ms host stall with torch loaded
        barrier()
        l0 = nat.kernel_launches()
        r0 = pipe.graphed.replays if pipe.graphed is not None else 0
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        t0.record()
        last = None
        marks = []
        for _ in range(steps):
            last = fn()
            pipe.collect_count_ms()
            marks.append(torch.cuda.Event(enable_timing=True))
            marks[-1].record()
        tail = pipe.flush()                          # pipelined e2e: the last batch's outputs are read inside the timed region
        last = tail if tail is not None else last
        if getattr(pipe, "up", None) is not None:   # the prefetch issued by the last step belongs to the timed region
            torch.cuda.current_stream().wait_stream(pipe.up.copy_stream)
        t1.record()
        barrier()
        gc.enable()
        wall = time.perf_counter() - w0
        ms = t0.elapsed_time(t1)
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        pipe.step_ms = [round(a.elapsed_time(b), 3) for a, b in zip([t0] + marks[:-1], marks)]   # diagnostic: per-step device time
        launches = nat.kernel_launches() - l0
        if pipe.graphed is not None:  # kernels replayed from the captured CUDA graph do not pass the host-side counter
            launches += (pipe.graphed.replays - r0) * pipe.graphed.kernels_per_replay
        return ms / steps, wall / steps, launches, last

    sampler = ClockSampler(local)
    sampler.start()
    ms_step, wall_step, launches, last = timed(pipe.step_resident, args.steps, args.warmup)
    clocks = sampler.stop()
    step_ms_resident = pipe.step_ms
    count_ms = statistics.mean(pipe.count_ms) if pipe.count_ms else None
    loss, emb, graph = last
    residues = NSEQ * SEQ_LEN * world
    # end to end: host bytes in, embeddings out
    e2e_ms, e2e_wall, _, last_e2e = timed(pipe.step_e2e, max(1, args.steps), max(args.warmup, 6))  # the e2e path warms its own allocations
    graph_h = last_e2e[2]
    h2d = int(pipe.h_buf.numel()) + 256 + 256      # the corpus as it crosses PCIe (5-bit symbols unless --host-bytes) + alphabet tables
    pat = graph_h.mathcal_A_out._nnz()
    d2h = (graph_h.number_of_nodes * 8 + 3 * graph_h.number_of_edges * 8 + 16  # node codes, edge table, sizes
           + graph_h.number_of_edges * (16 + 4) * 2 + pat * (16 + 12)          # A_out/A_in COO + pattern + 3 value arrays
           + graph_h.number_of_nodes * DIMS[-1] * 4 + 4 + 1024)                # embeddings, loss, alphabet table
    line = {
        "metric": METRIC, "value": residues / (ms_step * 1e-3), "unit": "residues/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8 keys / u64 counts / f32", "data": "synthetic",
        "config": step_config(
            nodes=graph.number_of_nodes, unique_edges=graph.number_of_edges, pattern_nnz=int(graph.mathcal_A_out._nnz()),
            l2_handling="inputs larger than L2 (corpus 176 MB per GPU > 126 MB L2)", cpu_affinity=affinity,
            multi_gpu="corpus sharded by sequence range, tables merged by NCCL all-reduce, n=3 DirectGCN replicated",
            directgcn_step="eager" if args.no_cuda_graph else "CUDA graph replay (fwd+loss+bwd+Adam+eval fwd captured once)",
            pipelining=("graph build of batch k+1 on its own stream under the DirectGCN replay of batch k (two-stage software pipeline; "
                        "e2e reads batch k's outputs while batch k+1 is built)") if pipe.pipelined else "none (sequential step)"),
        "clocks": clocks, "gpu_launches": int(launches / max(1, args.steps)),
        "e2e": {"value": residues / (e2e_ms * 1e-3), "unit": "residues/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h),
                "host_format": ("1 byte per symbol: the corpus buffer exactly as the FASTA reader leaves it in pinned host memory; nothing is "
                                "prepared outside the timed region") if pipe.h_packed is None else
                               "5-bit symbols (8 per 5 bytes, packed ONCE outside the timed region by pg_pack5_host), unpacked on the device (pg_unpack5)",
                "path": "host bytes -> H2D -> pg_ngram_count / extract / normalise (build_level_graph's kernels) -> DirectedNgramGraph on the host "
                        "(all five matrices + node names) -> ProtGramDirectGCN train step + extraction (CUDA-graph replay) -> embeddings + loss on the host"},
        "wall_ms_per_step": wall_step * 1e3,
        "per_step_ms": {"resident": step_ms_resident, "e2e": pipe.step_ms},
    }
    # the concurrent host -> device bound of this box: every rank uploads its corpus bytes from its pinned buffer at the same time
    # (what every e2e step must do); max over ranks.  e2e cannot beat it: on an 8-GPU node the ranks share the host's PCIe / memory system
    try:
        src = pipe.h_buf if torch.is_tensor(pipe.h_buf) else None
        if src is not None and src.is_pinned():
            dst = torch.empty(src.numel(), dtype=torch.uint8, device=dev)
            def _up():
                dst.copy_(src, non_blocking=True)
            barrier()
            bound_ms = _max_over_ranks(dist, dev, _time_ms(_up, 8, warm=2))
            line["e2e"]["h2d_concurrent_bound"] = {"ms_per_upload": bound_ms, "gbs_per_gpu": src.numel() / (bound_ms * 1e-3) / 1e9,
                                                   "gbs_aggregate": world * src.numel() / (bound_ms * 1e-3) / 1e9,
                                                   "e2e_step_over_bound": e2e_ms / bound_ms,
                                                   "note": "all ranks copy their step's corpus bytes pinned host -> HBM simultaneously, nothing else running"}
            del dst
    except Exception as exc:  # noqa: BLE001
        line["e2e"]["h2d_concurrent_bound"] = {"error": repr(exc)}
    if count_ms:
        alg = pipe.nbytes
        # the same call timed alone (no DirectGCN replay of the previous batch sharing the SMs)
        symbols_, d_rank_ = pipe.corpus.discover_alphabet(pipe.d_buf, None)
        ws_ = pipe.db.count_workspace(N_LEVEL, int(symbols_.size), pipe.nbytes, dev)
        pn_, pm_ = pipe.db.table_sizes(N_LEVEL, int(symbols_.size))
        bins_, short_ = torch.zeros(pm_, dtype=torch.int64, device=dev), torch.zeros(pn_, dtype=torch.uint8, device=dev)
        alone_ms = _time_ms(lambda: pipe.db.count_level(pipe.d_buf, N_LEVEL, d_rank_, int(symbols_.size), bins_, short_, ws_), 10, warm=3)
        tr_count = measured_traffic("ngram_count_smem_kernel", "r02_count_traffic.json")
        line["roofline_count"] = {"kernel": f"ngram_count_smem_kernel<M={N_LEVEL + 1}, 8-bit lanes> (+ memset, reduce_partials_kernel<8>, 2 gated no-op launches: "
                                      "everything pg_ngram_count enqueues, timed as one)",
                            "bound": "hbm", "achieved": alg / (count_ms * 1e-3) / 1e9,
                            "peak": peak_gbs, "unit": "GB/s", "frac": alg / (count_ms * 1e-3) / 1e9 / peak_gbs,
                            "traffic": tr_count["bytes_per_call"] if tr_count else None,
                            "traffic_source": tr_count["source"] if tr_count else "no round-2 capture committed (round 1: 181.3 MB, profiles/r01_ncu_count_smem8_v3.txt)",
                            "peak_source": peak_src, "algorithmic_bytes_per_launch": alg, "ms_per_launch": count_ms,
                            "share_of_step": count_ms / ms_step,
                            "ms_per_launch_alone": alone_ms, "frac_alone": alg / (alone_ms * 1e-3) / 1e9 / peak_gbs,
                            "shared_atomic_bound": {"achieved_Gupdates_per_s": NSEQ * SEQ_LEN / (alone_ms * 1e-3) / 1e9, "peak_Gupdates_per_s": 1300.0,
                                                    "peak_source": "tools/microbench_atomics.cu (shared-memory atomicAdd, 148 SMs)",
                                                    "note": "one shared-memory atomic per window is the floor of this kernel; includes memset + reduce + gated launches"},
                            "note": "ms_per_launch is measured inside the timed region, where the count kernel shares the SMs with the previous batch's "
                                    "DirectGCN replay (pipelining); ms_per_launch_alone is the same call without that overlap.  "
                                    "1 B/residue read (DRAM traffic == algorithmic bytes).  A 194,481-bin histogram cannot run at the HBM "
                                    "roofline: the count kernel alone takes 147 us under ncu, 88% issue-active (about 400 instructions per "
                                    "16 residues: rank lookup, rolling key, validity mask, packed 8-bit shared-memory add, overflow check); "
                                    "see DESIGN.md section 4.  The HBM-bound kernel of this system is the SpMM: `roofline`."}
        line["build"] = {"count_residues_per_s": NSEQ * SEQ_LEN / (count_ms * 1e-3)}
    if rank == 0:
        line["phases_ms"] = phase_breakdown(pipe)
    if rank == 0 and args.profile_host:
        import cProfile
        import pstats
        prof = cProfile.Profile()
        prof.enable()
        for _ in range(10):
            pipe.step_e2e() if args.profile_e2e else pipe.step_resident()
        torch.cuda.synchronize()
        prof.disable()
        with open(args.profile_host, "w") as fh:
            pstats.Stats(prof, stream=fh).sort_stats("cumulative").print_stats(45)
    roof_note = ("DirectGCN propagation kernel (north star: >= 60 % of the HBM roofline).  achieved = ALGORITHMIC bytes / CUDA-event time, "
                 "with SURVEY 8(d)'s model for operands beyond L2: one F-wide row fetch per stored entry (G = P).  The power-law graph "
                 "concentrates its entries on hub rows that stay in the 126 MB L2, so the DRAM traffic ncu measures (`traffic`) is BELOW "
                 "the algorithmic bytes; `frac_by_measured_traffic` is the fraction of the HBM peak the kernel really draws.  "
                 "roofline_count keeps the graph-build kernel of the headline step.")
    if rank == 0 and world == 1 and not args.no_large:
        try:
            leg = spmm_large_leg(pipe, peak_gbs, args.large_log2_nodes)
            line["spmm_large"] = leg
            fo = leg["fanout_fwd"]
            line["roofline"] = {"kernel": "spmm_fanout_kernel<3, 32, 1> (rows pass + long-row slices + reduce: everything pg_spmm_fanout enqueues, timed as one)",
                                "workload": leg["graph"] + f", F = {leg['F']}", "bound": "hbm", "achieved": fo["achieved_gbs"], "peak": peak_gbs,
                                "unit": "GB/s", "frac": fo["frac_of_hbm_peak"], "traffic": fo.get("dram_traffic_bytes_ncu"),
                                "traffic_source": fo.get("traffic_source"), "frac_by_measured_traffic": fo.get("frac_of_hbm_peak_by_ncu_traffic"),
                                "peak_source": peak_src, "algorithmic_bytes_per_launch": fo["algorithmic_bytes"], "ms_per_launch": fo["ms"],
                                "edges_per_s": fo["edges_per_s"],
                                "backward": {k: leg[k] for k in ("fanout_scaled_bwd", "fanin_bwd_operator")}, "note": roof_note}
        except Exception as exc:  # noqa: BLE001 - the headline must still print
            line["spmm_large"] = {"error": repr(exc)}
    if world > 1 and not args.no_large:
        try:
            leg = spmm_partitioned_leg(pipe, peak_gbs, dist, args.large_log2_nodes)
        except Exception as exc:  # noqa: BLE001 - the headline must still print
            leg = {"error": repr(exc)}
        if rank == 0:
            line["spmm_partitioned"] = leg
            if "fwd" in leg:
                f_ = leg["fwd"]
                line["roofline"] = {"kernel": "spmm_fanout_kernel<3, 32, 1> on this rank's row block over [own rows | halo rows] (pg_spmm_fanout_split), timed alone",
                                    "workload": leg["graph"] + f", F = {leg['F']}", "bound": "hbm", "achieved": f_["local_spmm_gbs"], "peak": peak_gbs,
                                    "unit": "GB/s", "frac": f_["local_spmm_frac_of_hbm_peak"], "traffic": None, "peak_source": peak_src,
                                    "algorithmic_bytes_per_launch": f_["local_spmm_algorithmic_bytes"], "ms_per_launch": f_["ms_local_spmm_alone"],
                                    "nvlink": {"exchange_ms": f_["ms_exchange_alone"], "gbs_this_gpu": f_["nvlink_gbs_this_gpu_exchange_alone"],
                                               "frac_of_measured_770": f_["nvlink_frac_of_measured_770"]}, "note": roof_note}
        # config C5 at its stated size (50 M-class nodes / 1 B edges, hidden 128, 8 GPUs): R-MAT needs a power of two -> 2^26 nodes, 1.07 G edges
        if world == 8 and not args.no_c5_full and args.large_log2_nodes < 23:
            torch.cuda.empty_cache()
            try:
                # NCCL transport here: at this size every byte of HBM headroom goes to the operands (the peer-memory ring of this
                # width alone is 30 GB); the weak-scaled leg above runs the peer-memory transport
                leg = spmm_partitioned_leg(pipe, peak_gbs, dist, 23, iters=3, compare_allgather=False, transport="nccl")
            except Exception as exc:  # noqa: BLE001
                leg = {"error": repr(exc)}
            if rank == 0:
                line["spmm_c5_full_size"] = leg
    if not args.no_scale:
        # BASELINE configs C3 (n=4) and C4 (n=5, 50 M sequences): every rank takes part (sharded corpus, NCCL merge)
        for key, n_level, seqs in (("build_c3_n4", 4, args.c3_seqs), ("build_c4_n5", 5, args.c4_seqs)):
            g = None
            try:
                leg = build_scale_leg(pipe, dist, n_level, seqs, normalise=(n_level == 4 or world == 1))
                g = leg.pop("_graph", None)
            except Exception as exc:  # noqa: BLE001 - the headline must still print
                leg = {"error": repr(exc)}
            if rank == 0:
                line[key] = leg
            if rank == 0 and n_level == 4 and g is not None:
                try:
                    line["directgcn_c3"] = c3_directgcn_leg(pipe, *g)
                except Exception as exc:  # noqa: BLE001
                    line["directgcn_c3"] = {"error": repr(exc)}
            del g
            torch.cuda.empty_cache()
            if dist is not None:
                dist.barrier()
    if rank == 0 and world == 1 and not args.no_large:
        try:
            line["layer_gemm_c3"] = layer_gemm_leg(pipe)
        except Exception as exc:  # noqa: BLE001
            line["layer_gemm_c3"] = {"error": repr(exc)}
        try:
            line["pooling_f2"] = pooling_leg(pipe, emb, graph)
        except Exception as exc:  # noqa: BLE001
            line["pooling_f2"] = {"error": repr(exc)}
    if rank == 0 and world == 1 and not args.no_large:
        try:
            line["e2e_api"] = e2e_api_leg(pipe)
        except Exception as exc:  # noqa: BLE001
            line["e2e_api"] = {"error": repr(exc)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_reference_pass(sample_seqs=6000, procs=1)
    if rank == 0:
        if "roofline" not in line and "roofline_count" in line:      # the SpMM leg was skipped (--no-large) or failed: keep the contract's key
            line["roofline"] = dict(line["roofline_count"], note="SpMM leg not run in this invocation: this is the count kernel's block (see roofline_count)")
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on host cores
# --------------------------------------------------------------------------------------------
def _count_chunk(args):
    first, nseq, n = args
    from oracle import ngram_oracle
    seqs = ngram_oracle.synth_sequences(first, nseq, SEQ_LEN, SEED)      # input generation: NOT timed
    padded = [(" " if first + i == 0 else "") + s + " " for i, s in enumerate(seqs)]
    import collections
    t0 = time.perf_counter()
    cnt = collections.Counter()
    grams = set()
    for p in padded:  # per-residue Python exactly like reference data_builder.py:38-54
        for i in range(len(p) - n + 1):
            grams.add(p[i:i + n])
        for i in range(len(p) - n):
            cnt[(p[i:i + n], p[i + 1:i + 1 + n])] += 1
    return grams, cnt, time.perf_counter() - t0


def cpu_reference_pass(sample_seqs: int, procs: int):
    """One bounded pass of the reference algorithm on host cores.  Builder: `sample_seqs` of the 500k
    sequences (per-residue Python, `procs` processes; the synthetic sequences are generated before the clock
    starts); graph + DirectGCN: the full C2-sized graph (a 6k-sequence sample already saturates the 8.4k 3-grams).
    value = full-workload residues/s extrapolated as full/(t_build*full/sample + t_graph + t_gcn)."""
    import collections
    from oracle import directgcn_oracle, graph_oracle
    per = max(1, sample_seqs // procs)
    jobs = [(i * per, per, N_LEVEL) for i in range(procs)]
    if procs > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(procs) as pool:
            parts = pool.map(_count_chunk, jobs)
    else:
        parts = [_count_chunk(j) for j in jobs]
    t0 = time.perf_counter()
    grams, cnt = set(), collections.Counter()
    for g_, c_, _ in parts:
        grams |= g_
        cnt.update(c_)
    nodes = sorted(grams)
    ids = {g_: i for i, g_ in enumerate(nodes)}
    keys = sorted((ids[a], ids[b]) for a, b in cnt)
    inv = {(ids[a], ids[b]): c for (a, b), c in cnt.items()}
    src = np.array([k[0] for k in keys], dtype=np.int64)
    dst = np.array([k[1] for k in keys], dtype=np.int64)
    w = np.array([inv[k] for k in keys], dtype=np.int64)
    t_build = max(p_[2] for p_ in parts) + (time.perf_counter() - t0)     # slowest worker's counting loop + the merge
    t1 = time.perf_counter()
    n = len(nodes)
    mats = graph_oracle.normalise_all(src, dst, w, n)
    t_graph = time.perf_counter() - t1
    torch.set_num_threads(os.cpu_count() or 1)
    params = directgcn_oracle.init_params(DIMS, n, n, seed=SEED)
    opt = torch.optim.Adam(list(params.values()), lr=LR)
    x = torch.randn(n, DIMS[0])
    e = lambda m: (torch.from_numpy(np.stack([mats[m][0], mats[m][1]])), torch.from_numpy(mats[m][2]))
    (ei_in, ew_in), (ei_out, ew_out), (ei_un, ew_un) = e("mathcal_A_in"), e("mathcal_A_out"), e("A_undirected_norm_sparse")
    a_out = torch.sparse_coo_tensor(torch.from_numpy(np.stack([src, dst])), torch.from_numpy(w.astype(np.float32)), (n, n)).coalesce()
    labels = next_node_labels(a_out, n)
    t2 = time.perf_counter()
    opt.zero_grad()
    logp, _ = directgcn_oracle.protgram_forward(params, x, ei_in, ew_in, ei_out, ew_out, ei_un, ew_un, N_LEVEL, 0)
    loss = torch.nn.functional.nll_loss(logp, labels) + L2_LAMBDA * sum(p.norm(2).pow(2) for p in params.values())
    loss.backward()
    opt.step()
    with torch.no_grad():
        directgcn_oracle.protgram_forward(params, x, ei_in, ew_in, ei_out, ew_out, ei_un, ew_un, N_LEVEL, 0)
    t_gcn = time.perf_counter() - t2
    full = NSEQ * SEQ_LEN
    sample = per * procs * SEQ_LEN
    t_full = t_build * full / sample + t_graph + t_gcn
    return {"value": full / t_full, "unit": "residues/s", "cores": max(procs, 1), "kind": "port",
            "torch_threads_for_directgcn": torch.get_num_threads(),
            "sample": f"builder: {per * procs} of {NSEQ} sequences ({sample} residues) in {t_build:.2f}s with {procs} process(es) of per-residue "
                      f"Python (oracle restatement of data_builder.py:38-54, no Dask/text/CSV overhead; sequence generation not timed); graph "
                      f"normalisation {t_graph:.2f}s and DirectGCN train step + extraction {t_gcn:.2f}s on the full-size graph ({n} nodes, "
                      f"{len(keys)} edges); value = {full} / (t_build*{full / sample:.1f} + t_graph + t_gcn)",
            "extrapolation_factor_builder": full / sample, "nodes": n, "unique_edges": len(keys),
            "t_build_sample_s": t_build, "t_graph_s": t_graph, "t_gcn_s": t_gcn, "t_measured_s": t_build + t_graph + t_gcn,
            "builder_residues_per_s": sample / t_build}


def step_config(**over):
    """The `config` object both arms print (same keys; what an arm does not have stays None)."""
    cfg = {"workload": WORKLOAD, "residues_per_step_per_gpu": NSEQ * SEQ_LEN, "n": N_LEVEL, "layer_dims": DIMS, "nodes": None,
           "unique_edges": None, "pattern_nnz": None, "l2_handling": None, "cpu_affinity": None, "multi_gpu": None, "directgcn_step": None,
           "pipelining": None, "note": None}
    unknown = set(over) - set(cfg)
    assert not unknown, unknown
    cfg.update(over)
    return cfg


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return
    procs = os.cpu_count() or 1
    sample = max(2000, 1500 * procs)
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_reference_pass(min(sample, 2000), procs)
    res, t0 = [], time.perf_counter()
    for _ in range(args.steps):
        res.append(cpu_reference_pass(sample, procs))
    wall = (time.perf_counter() - t0) / args.steps
    single = cpu_reference_pass(3000, 1)     # the reference's effective mode (to_textfiles is forced `sync`, the Bag ops are GIL-bound threads)
    best = max(res, key=lambda r: r["value"])
    value = statistics.mean(r["value"] for r in res)
    best["value"] = value
    best["single_process"] = {k: single[k] for k in ("value", "cores", "sample", "builder_residues_per_s")}
    measured_ms = statistics.mean(r["t_measured_s"] for r in res) * 1e3
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "residues/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": measured_ms, "ms_per_full_step_extrapolated": NSEQ * SEQ_LEN / value * 1e3,
        "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "python str / int / f32", "data": "synthetic",
        "config": step_config(nodes=best["nodes"], unique_edges=best["unique_edges"],
                              note=f"reference algorithm (oracle port, kind=port) on {procs} host processes; each step = a bounded sample "
                                   f"({sample} of {NSEQ} sequences through the per-residue builder, extrapolation factor "
                                   f"{best['extrapolation_factor_builder']:.1f} on the builder time only) + normalisation and DirectGCN step on the "
                                   f"full-size graph; ms_per_step is the MEASURED time of one sample step, value the full-workload throughput it "
                                   f"extrapolates to; wall per sample step incl. untimed sequence generation {wall:.1f} s"),
        "cpu_baseline": best, "gpu_launches": 0,
        "e2e": {"value": value, "unit": "residues/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-large", action="store_true", help="skip the large-graph SpMM leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--host-packed5", action="store_true", help="e2e: start from a corpus that the ingest already keeps as 5-bit symbols in host "
                    "memory (packed once, outside the timed region) instead of 1 byte per symbol")
    ap.add_argument("--no-pipeline", action="store_true", help="strictly sequential steps (no overlap of batch k+1's build with batch k's DirectGCN step)")
    ap.add_argument("--no-cuda-graph", action="store_true", help="run the DirectGCN step eagerly instead of replaying the captured CUDA graph")
    ap.add_argument("--large-log2-nodes", type=int, default=21)
    ap.add_argument("--no-c5-full", action="store_true", help="8 GPUs: skip config C5 at its full size (2^26 nodes, 1.07 G edges)")
    ap.add_argument("--no-scale", action="store_true", help="skip the C3 (n=4) / C4 (n=5) build legs and the C3 DirectGCN leg")
    ap.add_argument("--c3-seqs", type=int, default=2_000_000, help="sequences of the n=4 build leg (whole job)")
    ap.add_argument("--c4-seqs", type=int, default=50_000_000, help="sequences of the n=5 build leg (whole job; 17.5 G residues)")
    ap.add_argument("--profile-host", default=None, help="write a cProfile of 10 resident steps to this file")
    ap.add_argument("--profile-e2e", action="store_true", help="profile e2e steps instead of resident ones")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
