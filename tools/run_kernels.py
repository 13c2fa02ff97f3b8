#!/usr/bin/env python
"""Runs the two dominant kernels a few times at benchmark size, for `ncu --set full` captures:
   ngram_count (C2 corpus, n=3) and the nv=3 fan-out / fan-in SpMM on the R-MAT graph of bench.py.
usage: python tools/run_kernels.py [count] [count4] [count5] [spmm] [--log2 N] [--seqs S]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["count", "spmm"]
    log2 = int(sys.argv[sys.argv.index("--log2") + 1]) if "--log2" in sys.argv else 21
    torch.cuda.set_device(0)
    pipe = bench.B200Pipeline(0, 1, torch.device("cuda", 0))
    if "count" in which:
        symbols, d_rank = pipe.corpus.discover_alphabet(pipe.d_buf)
        for _ in range(3):
            bins, short = pipe.db.count_level(pipe.d_buf, bench.N_LEVEL, d_rank, int(symbols.size))
        torch.cuda.synchronize()
        print("count ok", int(bins.sum()))
    for key, n_level in (("count4", 4), ("count5", 5)):   # variant P (partition + shared-memory count) at C3 / C4 shape
        if key in which:
            seqs = int(sys.argv[sys.argv.index("--seqs") + 1]) if "--seqs" in sys.argv else 2_000_000
            out = bench.build_scale_leg(pipe, None, n_level, seqs, iters=1, normalise=False)
            print({k: v for k, v in out.items() if not k.startswith("_")})
    if "c3step" in which:   # config C3's DirectGCN step (3 layers, hidden 256) on the n=4 graph
        out = bench.build_scale_leg(pipe, None, 4, 2_000_000, iters=1, normalise=True)
        nodes, res = out.pop("_graph")
        print(bench.c3_directgcn_leg(pipe, nodes, res, iters=1))
    if "spmm" in which:
        out = bench.spmm_large_leg(pipe, 6548.5, log2, iters=2)
        print({k: (v if not isinstance(v, dict) else {a: round(b, 3) for a, b in v.items()}) for k, v in out.items()})


if __name__ == "__main__":
    main()
