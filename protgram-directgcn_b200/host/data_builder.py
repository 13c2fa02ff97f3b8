"""GraphBuilder -- drop-in for the reference's src/pipeline/data_builder.py:57-341.

Same constructor (`GraphBuilder(config)`), same `run()` contract: for n = 1..GCN_NGRAM_MAX_N write
`GRAPH_OBJECTS_DIR/ngram_graph_n{n}.pkl` holding a DirectedNgramGraph; print-and-continue on data
errors, never raise from run() for them.  What changed is the body: the Dask bag / text spill /
CSV re-parse / groupby pipeline (reference :118-220,267-273) is one pass of the count kernel over
the corpus resident in HBM plus a scan/compaction (csrc/ngram.cu), and the graph object is built
from the device edge table without the parquet round trip.

Multi-GPU (one process per GPU, torch.distributed): each rank packs and counts a contiguous
slice of the sequences; the dense tables are summed with an all-reduce (the merge step of
SURVEY.md 8(e)); every rank then extracts the identical graph and rank 0 writes the pickles.
`build_level_graph_partitioned` is the variant for graphs that should never exist on one GPU:
reduce-scatter of the tables over key ranges, per-range extraction, row-partitioned normalisation.
"""
from __future__ import annotations

import os
import time
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from .. import _native as nat
from . import corpus
from .data_utils import DataLoader, DataUtils
from .graph_utils import DirectedNgramGraph


def table_sizes(n: int, sigma: int) -> Tuple[int, int]:
    return sigma ** n, sigma ** (n + 1)


def count_workspace(n: int, sigma: int, nbytes: int, device) -> torch.Tensor:
    """Workspace of pg_ngram_count for corpus buffers of up to `nbytes` (reusable across chunks and calls)."""
    return nat.workspace(nat.query("pg_ngram_count_ws_bytes_for", n, sigma, int(nbytes)), device)


def count_level(d_buf: torch.Tensor, n: int, d_rank: torch.Tensor, sigma: int,
                bins: Optional[torch.Tensor] = None, short: Optional[torch.Tensor] = None,
                ws: Optional[torch.Tensor] = None):
    """Accumulate the (n+1)-gram table of one corpus buffer.  -> (bins int64[sigma^(n+1)],
    short_present uint8[sigma^n]).  Replaces reference data_builder.py:45-54,203-220,267-273.
    ws: optional workspace from count_workspace() (multi-GB for n >= 4: allocate it once per level)."""
    nat.require_cuda()
    pow_n, pow_m = table_sizes(n, sigma)
    dev = d_buf.device
    if bins is None:
        bins = torch.zeros(pow_m, dtype=torch.int64, device=dev)
    if short is None:
        short = torch.zeros(pow_n, dtype=torch.uint8, device=dev)
    if ws is None:
        ws = count_workspace(n, sigma, d_buf.numel(), dev)
    nat.call("pg_ngram_count", nat.ptr(d_buf), d_buf.numel(), n, nat.ptr(d_rank), sigma, nat.ptr(bins), nat.ptr(short),
             nat.ptr(ws), ws.numel(), nat.stream_ptr())
    return bins, short


def extract_level(bins: torch.Tensor, short: torch.Tensor, n: int, sigma: int):
    """Dense table -> (node_code int64[N] ascending, src, dst, count int64[E] sorted by (src,dst)).
    Replaces reference data_builder.py:151-177 (distinct + sorted ids) and :281-286."""
    dev = bins.device
    ws = nat.workspace(nat.query("pg_graph_extract_ws_bytes", n, sigma), dev)
    sizes = torch.zeros(2, dtype=torch.int64, device=dev)
    st = nat.stream_ptr()
    nat.call("pg_graph_extract_sizes", nat.ptr(bins), nat.ptr(short), n, sigma, nat.ptr(sizes), nat.ptr(ws), ws.numel(), st)
    num_nodes, num_edges = (int(v) for v in sizes.tolist())
    node_code = torch.empty(num_nodes, dtype=torch.int64, device=dev)
    src = torch.empty(num_edges, dtype=torch.int64, device=dev)
    dst = torch.empty(num_edges, dtype=torch.int64, device=dev)
    cnt = torch.empty(num_edges, dtype=torch.int64, device=dev)
    nat.call("pg_graph_extract_fill", nat.ptr(bins), n, sigma, num_nodes, num_edges, nat.ptr(node_code), nat.ptr(src),
             nat.ptr(dst), nat.ptr(cnt), nat.ptr(ws), ws.numel(), st)
    return node_code, src, dst, cnt


def build_level_graph(d_buf, n: int, symbols: np.ndarray, d_rank: torch.Tensor, eps: float,
                      group=None) -> DirectedNgramGraph:
    """corpus buffer(s) -> DirectedNgramGraph for one n (count, merge, extract, normalise).
    d_buf: a device corpus buffer, or a list of chunks (device tensors or host arrays that are
    uploaded one at a time: corpora larger than HBM stream through, the tables accumulate)."""
    sigma = int(symbols.size)
    chunks = list(d_buf) if isinstance(d_buf, (list, tuple)) else [d_buf]
    bins = short = None
    biggest = max((int(c.numel()) if torch.is_tensor(c) else int(c.size) for c in chunks), default=0)   # PackedChunk.size = logical bytes
    ws = count_workspace(n, sigma, biggest, d_rank.device) if chunks else None   # one workspace for all chunks of the level
    host_idx = [i for i, c in enumerate(chunks) if not (torch.is_tensor(c) and c.is_cuda)]
    up = corpus.CorpusUploader(d_rank.device) if host_idx else None
    wire = lambda c: c.packed if isinstance(c, corpus.PackedChunk) else c     # what crosses PCIe: 5-bit symbols when the chunk was packed
    scratch = None
    if up is not None:
        up.submit(wire(chunks[host_idx[0]]))
    for i, c in enumerate(chunks):
        if torch.is_tensor(c) and c.is_cuda:
            bins, short = count_level(c, n, d_rank, sigma, bins, short, ws)
            continue
        nxt = host_idx.index(i) + 1
        if nxt < len(host_idx):
            up.submit(wire(chunks[host_idx[nxt]]))      # upload of the next chunk runs under this chunk's count
        d_chunk = up.acquire()
        if isinstance(c, corpus.PackedChunk):
            if scratch is None or scratch.numel() < c.n_symbols:
                scratch = torch.empty(max(c.n_symbols, 16), dtype=torch.uint8, device=d_rank.device)
            d_chunk = corpus.unpack5(d_chunk, c.n_symbols, scratch)
        bins, short = count_level(d_chunk, n, d_rank, sigma, bins, short, ws)
        up.release()
    if bins is None:
        bins, short = count_level(torch.empty(0, dtype=torch.uint8, device=d_rank.device), n, d_rank, sigma)
    if group is not None and _use_key_range_merge(n, sigma, group):
        return _finish_by_key_range(bins, short, n, symbols, sigma, eps, group)
    if group is not None:
        import torch.distributed as dist
        dist.all_reduce(bins, op=dist.ReduceOp.SUM, group=group)
        short_i = short.to(torch.int32)
        dist.all_reduce(short_i, op=dist.ReduceOp.MAX, group=group)
        short = short_i.to(torch.uint8)
    node_code, src, dst, cnt = extract_level(bins, short, n, sigma)
    names = corpus.LazyNodeNames(node_code, symbols, n)   # id -> n-gram string, id order; decoded on first access
    return DirectedNgramGraph.from_edge_arrays(names, src, dst, cnt.to(torch.float32), epsilon_propagation=eps,
                                               n_value=n, assume_coalesced=True)


KEY_RANGE_MERGE_MIN_TABLE_BYTES = 126 << 20     # B200's L2: beyond it the all-reduced table is re-read from HBM by every rank's extraction


def _use_key_range_merge(n: int, sigma: int, group) -> bool:
    """Multi-GPU merge policy of one level: tables that fit L2 (n <= 4 at sigma = 21: 33 MB) are all-reduced and every rank
    extracts the identical graph; larger ones (n = 5: 686 MB) are REDUCE-SCATTERED over key ranges, every rank extracts the
    edges of its own range only and the edge lists (16 B per edge, not 8 B per bin) are gathered on rank 0 for the pickle."""
    import torch.distributed as dist
    return dist.get_world_size(group) > 1 and table_sizes(n, sigma)[1] * 8 > KEY_RANGE_MERGE_MIN_TABLE_BYTES


def gather_edges_on_root(src, dst, cnt, group, root: int = 0):
    """Variable-size gather of the per-rank edge lists (already globally sorted in rank order: key ranges ascend) onto `root`
    -- one all_to_all_single in which only the root receives.  -> (src, dst, cnt) on root, three empty tensors elsewhere."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = src.device
    sizes = torch.zeros(world, dtype=torch.int64, device=dev)
    sizes[rank] = src.numel()
    dist.all_reduce(sizes, group=group)
    sizes = [int(v) for v in sizes.tolist()]
    send = torch.stack([src, dst, cnt], dim=1).contiguous()                       # [e, 3] int64
    in_split = [send.shape[0] if r == root else 0 for r in range(world)]
    out_split = sizes if rank == root else [0] * world
    recv = torch.empty((sum(out_split), 3), dtype=torch.int64, device=dev)
    dist.all_to_all_single(recv, send, output_split_sizes=out_split, input_split_sizes=in_split, group=group)
    return recv[:, 0].contiguous(), recv[:, 1].contiguous(), recv[:, 2].contiguous()


def _finish_by_key_range(bins, short, n: int, symbols, sigma: int, eps: float, group):
    """count tables -> reduce-scatter over key ranges -> per-range extraction -> edge lists gathered on rank 0, which alone
    builds (and later pickles) the whole-graph object; the other ranks return a node-list-only stub (`edges_on_rank0`)."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    pow_n, pow_m = table_sizes(n, sigma)
    codes_per = (pow_n + world - 1) // world
    padded = torch.zeros(world * codes_per * sigma, dtype=torch.int64, device=bins.device)
    padded[:pow_m] = bins
    del bins
    local = merge_tables_by_key_range(padded, codes_per, sigma, group)
    del padded
    short_i = short.to(torch.int32)
    dist.all_reduce(short_i, op=dist.ReduceOp.MAX, group=group)
    node_code, src, dst, cnt = extract_key_range(local, short_i.to(torch.uint8), n, sigma, rank * codes_per, codes_per, group)
    total = torch.tensor([int(src.numel())], dtype=torch.int64, device=src.device)
    dist.all_reduce(total, group=group)
    src, dst, cnt = gather_edges_on_root(src, dst, cnt, group)
    names = corpus.LazyNodeNames(node_code, symbols, n)
    if rank == 0:
        return DirectedNgramGraph.from_edge_arrays(names, src, dst, cnt.to(torch.float32), epsilon_propagation=eps, n_value=n,
                                                   assume_coalesced=True)
    return _RemoteLevelGraph(n, names, int(node_code.numel()), int(total.item()))


class _RemoteLevelGraph:
    """What the non-root ranks of a key-range merged level hold: the (replicated) node list and the totals; the edges and
    matrices live on rank 0 (`GraphBuilder.run` only pickles there)."""
    edges_on_rank0 = True

    def __init__(self, n_value, names, number_of_nodes, number_of_edges):
        self.n_value, self._names = n_value, names
        self.number_of_nodes, self.number_of_edges = number_of_nodes, number_of_edges

    @property
    def node_sequences(self):
        return self._names.resolve()


class PartitionedLevelGraph:
    """One n level of a graph that is never assembled on one GPU (what `build_level_graph_partitioned`
    returns on every rank): the node list is replicated (ids need the global presence table), edges
    and matrices exist as this rank's row block [lo, hi) only.

    node_code       int64[N]   packed n-gram codes, ascending = id order (replicated)
    a_out           (src, dst, count) of the block's rows of A_out_w, sorted by (src, dst)
    block           dict from partitioned.normalize_row_partitioned (rowptr padded to per + 1, global
                    int32 columns, val_in / val_out / val_und, the block's rows of A_in_w)"""

    def __init__(self, n_value, symbols, node_code, a_out, block, number_of_edges, group):
        self.n_value, self.symbols, self.node_code, self.a_out, self.block = n_value, symbols, node_code, a_out, block
        self.number_of_nodes = int(node_code.numel())
        self.number_of_edges = number_of_edges
        self.lo, self.hi, self.per = block["lo"], block["hi"], block["per"]
        self.group = group

    @property
    def node_sequences(self):
        return corpus.LazyNodeNames(self.node_code, self.symbols, self.n_value).resolve()

    def propagation(self):
        """Row-partitioned SpMM over the block (host/partitioned.py)."""
        from . import partitioned as part
        return part.RowPartitionedPropagation.from_local(part.local_csr(self.block), self.number_of_nodes, group=self.group)


def merge_tables_by_key_range(bins_padded: torch.Tensor, codes_per: int, sigma: int, group) -> torch.Tensor:
    """Sum the per-rank tables and leave every rank with ITS key range only: rank r gets the bins of the source codes
    [r * codes_per, (r + 1) * codes_per) -- the north star's reduce-scatter over key ranges (half the bytes of an
    all-reduce, and no rank ever holds the merged table)."""
    import torch.distributed as dist
    rank = dist.get_rank(group)
    chunk = codes_per * sigma
    if dist.get_backend(group) == "nccl":
        local = torch.empty(chunk, dtype=bins_padded.dtype, device=bins_padded.device)
        dist.reduce_scatter_tensor(local, bins_padded, op=dist.ReduceOp.SUM, group=group)
        return local
    dist.all_reduce(bins_padded, op=dist.ReduceOp.SUM, group=group)   # gloo (CPU tests) has no reduce_scatter
    return bins_padded[rank * chunk:(rank + 1) * chunk].clone()


def extract_key_range(local_bins: torch.Tensor, short: torch.Tensor, n: int, sigma: int, code_lo: int, codes: int, group):
    """This rank's key range -> (node_code [N] replicated, src, dst, count of the range's edges with GLOBAL node ids).
    One small collective: the presence table (sigma^n bytes) is MAX-reduced so every rank numbers the nodes alike."""
    import torch.distributed as dist
    dev = local_bins.device
    pow_n = sigma ** n
    st = nat.stream_ptr()
    present = short.clone()
    ws = nat.workspace(nat.query("pg_graph_extract_range_ws_bytes", sigma, codes), dev)
    sizes = torch.zeros(2, dtype=torch.int64, device=dev)
    nat.call("pg_graph_extract_range_mark", nat.ptr(local_bins), n, sigma, code_lo, codes, nat.ptr(present), nat.ptr(sizes),
             nat.ptr(ws), ws.numel(), st)
    present_i = present.to(torch.int32)
    dist.all_reduce(present_i, op=dist.ReduceOp.MAX, group=group)
    present = present_i.to(torch.uint8)
    node_id = torch.empty(pow_n, dtype=torch.int64, device=dev)
    ws_ids = nat.workspace(nat.query("pg_node_ids_ws_bytes", pow_n), dev)
    nat.call("pg_node_ids_from_presence", nat.ptr(present), pow_n, nat.ptr(node_id), nat.ptr(sizes[1:]), nat.ptr(ws_ids), ws_ids.numel(), st)
    num_edges, num_nodes = (int(v) for v in sizes.tolist())
    node_code = torch.empty(num_nodes, dtype=torch.int64, device=dev)
    if num_nodes:
        nat.call("pg_node_codes_emit", nat.ptr(present), nat.ptr(node_id), pow_n, nat.ptr(node_code), st)
    src = torch.empty(num_edges, dtype=torch.int64, device=dev)
    dst = torch.empty(num_edges, dtype=torch.int64, device=dev)
    cnt = torch.empty(num_edges, dtype=torch.int64, device=dev)
    nat.call("pg_graph_extract_range_fill", nat.ptr(local_bins), n, sigma, code_lo, codes, nat.ptr(node_id), num_edges, nat.ptr(src),
             nat.ptr(dst), nat.ptr(cnt), nat.ptr(ws), ws.numel(), st)
    return node_code, src, dst, cnt


def build_level_graph_partitioned(d_buf: torch.Tensor, n: int, symbols: np.ndarray, d_rank: torch.Tensor, eps: float,
                                  group) -> PartitionedLevelGraph:
    """The fully partitioned build of one n level (SURVEY.md 8(e), rows "Builder" + "Normalisation"): every rank counts
    its corpus shard, the tables are merged by REDUCE-SCATTER over key ranges, every rank extracts the edges of its key
    range (= whole source rows), the edge lists are re-dealt onto equal row blocks and normalised with one exchange.
    Neither the merged table nor the graph ever exists on one GPU; results are bit-identical to `build_level_graph`."""
    import torch.distributed as dist
    from . import partitioned as part
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    sigma = int(symbols.size)
    pow_n, pow_m = table_sizes(n, sigma)
    codes_per = (pow_n + world - 1) // world
    dev = d_rank.device
    bins_padded = torch.zeros(world * codes_per * sigma, dtype=torch.int64, device=dev)
    _, short = count_level(d_buf, n, d_rank, sigma, bins=bins_padded[:pow_m])
    local_bins = merge_tables_by_key_range(bins_padded, codes_per, sigma, group)
    del bins_padded
    short_i = short.to(torch.int32)
    dist.all_reduce(short_i, op=dist.ReduceOp.MAX, group=group)
    node_code, src, dst, cnt = extract_key_range(local_bins, short_i.to(torch.uint8), n, sigma, rank * codes_per, codes_per, group)
    num_nodes = int(node_code.numel())
    if num_nodes == 0:
        raise ValueError(f"no n-grams for n={n}")
    total_edges = torch.tensor([int(src.numel())], dtype=torch.int64, device=dev)
    dist.all_reduce(total_edges, group=group)
    # key ranges are equal slices of the CODE space; the normalisation wants equal slices of the ID space
    src, dst, w = part.exchange_in_edges(src, dst, cnt.to(torch.float32), num_nodes, group, by="src")
    block = part.normalize_row_partitioned(src, dst, w, num_nodes, eps, group)
    return PartitionedLevelGraph(n, symbols, node_code, (src, dst, w), block, int(total_edges.item()), group)


class GraphBuilder:
    def __init__(self, config):
        self.config = config
        self.protein_sequence_file = str(config.GCN_INPUT_FASTA_PATH)
        self.output_dir = str(config.GRAPH_OBJECTS_DIR)
        self.n_max = config.GCN_NGRAM_MAX_N
        self.num_workers_config = config.GRAPH_BUILDER_WORKERS if config.GRAPH_BUILDER_WORKERS is not None else 1
        self.temp_dir = os.path.join(str(config.BASE_OUTPUT_DIR), "temp_graph_builder")
        self.gcn_propagation_epsilon = getattr(config, "GCN_PROPAGATION_EPSILON", 1e-9)
        self.process_group = getattr(config, "GRAPH_BUILDER_PROCESS_GROUP", None)  # optional, multi-GPU
        print(f"GraphBuilder initialized: n_max={self.n_max}, output_dir='{self.output_dir}' (CUDA path)")

    def _rank_world(self) -> Tuple[int, int]:
        if self.process_group is None:
            return 0, 1
        import torch.distributed as dist
        return dist.get_rank(self.process_group), dist.get_world_size(self.process_group)

    def _agree(self, flag: int, count: int) -> Tuple[int, int]:
        """-> (max of `flag`, sum of `count`) over the ranks of the process group (identity without one)."""
        if self.process_group is None:
            return flag, count
        import torch.distributed as dist
        dev = nat.current_device() if dist.get_backend(self.process_group) == "nccl" else torch.device("cpu")
        t = torch.tensor([flag, count], dtype=torch.int64, device=dev)
        f = t[:1].clone()
        dist.all_reduce(f, op=dist.ReduceOp.MAX, group=self.process_group)
        c = t[1:].clone()
        dist.all_reduce(c, op=dist.ReduceOp.SUM, group=self.process_group)
        return int(f.item()), int(c.item())

    def run(self):
        t0 = time.monotonic()
        DataUtils.print_header("PIPELINE STEP 1: Building N-gram Graphs")
        nat.load()
        nat.require_cuda()  # fail loudly: no CPU fallback
        os.makedirs(self.output_dir, exist_ok=True)
        if not os.path.exists(os.path.normpath(self.protein_sequence_file)):
            print(f"ERROR: FASTA file not found at {self.protein_sequence_file}")
            return
        rank, world = self._rank_world()
        dev = nat.current_device()
        chunk_bytes = int(getattr(self.config, "GRAPH_BUILDER_CHUNK_BYTES", 1 << 28))
        hbm_budget = int(getattr(self.config, "GRAPH_BUILDER_RESIDENT_BYTES", 0.5 * torch.cuda.mem_get_info()[0] if torch.cuda.is_available() else 1 << 62))
        n_seqs = [0]

        def sequences():
            for _, s in DataLoader.parse_sequences(self.protein_sequence_file):
                n_seqs[0] += 1
                yield s

        def host_chunks():
            # FASTA -> corpus buffer natively (csrc/fasta.cu: mmap, no Python string per sequence; row f3).  Non-ASCII
            # sequence text takes the Python parser, which reproduces the reference's utf-8 'ignore' decoding first.
            stats = {}
            try:
                path = os.path.normpath(self.protein_sequence_file)
                try:
                    ram = os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES")
                except (ValueError, OSError):
                    ram = 0
                if os.path.getsize(path) <= ram // 4:
                    # fits host memory comfortably: parse it with all host threads in one call, then cut at sequence boundaries
                    whole = corpus.read_fasta_parallel(path, rank=rank, world=world, pinned=True, stats=stats)
                    for buf in corpus.split_at_separators(whole, chunk_bytes):
                        yield buf
                else:
                    # larger than that: window by window, each window parsed by all host threads
                    window = int(getattr(self.config, "GRAPH_BUILDER_FASTA_WINDOW_BYTES", max(chunk_bytes, 1 << 30)))
                    for whole in corpus.stream_fasta_windows(path, window, rank=rank, world=world, pinned=True, stats=stats):
                        for buf in corpus.split_at_separators(whole, chunk_bytes):
                            yield buf
                n_seqs[0] = stats.get("sequences", 0)
                if stats.get("stopped_early"):
                    print(f"Error parsing FASTA file {self.protein_sequence_file}: list index out of range")  # the reference's message
                return
            except corpus.NonAsciiSequence:
                if chunks:
                    raise ValueError("FASTA holds non-ASCII sequence text; unsupported on the CUDA path")
            yield from corpus.stream_chunks(sequences(), chunk_bytes, rank, world)

        chunks, resident = [], 0
        failed = None
        try:
            for buf in host_chunks():
                size = int(buf.numel()) if torch.is_tensor(buf) else int(buf.size)
                if resident + size <= hbm_budget:          # keep the corpus in HBM across the n levels
                    chunks.append(corpus.to_device(buf, dev))
                    resident += size
                else:                                       # larger than the budget: re-streamed per level, 5 bits per symbol on the wire
                    chunks.append(corpus.pack5(buf) or buf)
        except ValueError as exc:
            failed = str(exc)
        # ranks agree on failure BEFORE the first collective: a rank that returned alone would leave the others blocked in it
        any_failed, total_seqs = self._agree(1 if failed else 0, n_seqs[0])
        if failed:
            print(f"ERROR: {failed}")
        if any_failed:
            if not failed:
                print("ERROR: another rank could not read its shard of the FASTA file. Cannot proceed.")
            return
        if total_seqs == 0:
            print("ERROR: No sequences found in the FASTA file. Cannot proceed.")
            return
        print(f"  Loaded {n_seqs[0]} sequences from FASTA ({len(chunks)} corpus chunk(s) on this rank).")
        try:
            symbols, d_rank = corpus.discover_alphabet(chunks, self.process_group)
        except ValueError as exc:          # raised from the all-reduced presence table: every rank sees the same error
            print(f"ERROR: {exc}")
            return
        d_buf = chunks
        for n in range(1, self.n_max + 1):
            t_level = time.monotonic()
            try:
                graph = build_level_graph(d_buf, n, symbols, d_rank, self.gcn_propagation_epsilon, self.process_group)
            except nat.NativeError:
                raise
            except Exception as exc:  # noqa: BLE001 - reference prints and continues with the next level
                print(f"  [n={n}] ERROR building graph: {exc}")
                continue
            if graph.number_of_nodes == 0:
                print(f"  Info: no n-grams for n={n}. No graph will be generated. Skipping.")
                continue
            if rank == 0:
                out_path = os.path.join(self.output_dir, f"ngram_graph_n{n}.pkl")
                DataUtils.save_object(graph, out_path)
                print(f"  Graph for n={n} saved to {out_path}")
            print(f"    Nodes: {graph.number_of_nodes}  Edges (unique weighted): {graph.number_of_edges}  "
                  f"[{time.monotonic() - t_level:.3f}s]")
        DataUtils.print_header(f"N-gram Graph Building FINISHED in {time.monotonic() - t0:.2f}s")
