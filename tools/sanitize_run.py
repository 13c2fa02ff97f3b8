#!/usr/bin/env python
"""Small-input pass over the library's kernels for `compute-sanitizer` (memcheck / racecheck batches).
ONE tool per gpurun call (B200_PROFILING.md):

    compute-sanitizer --tool memcheck  --error-exitcode 3 python tools/sanitize_run.py > gpurun_out/memcheck.log 2>&1
    compute-sanitizer --tool racecheck --error-exitcode 3 python tools/sanitize_run.py [groups] > gpurun_out/racecheck.log 2>&1

What runs: __graft_entry__.smoke() (count -> extract -> normalise -> DirectGCN forward / backward on the SIMT and the
tcgen05 paths, decoder, loss) and a handful of the parity tests at their smallest parameters (every count variant incl. the
lane-overflow protocol, radix sort, split-operand SpMM with long rows, row-block normalisation, next-level kernels)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    groups = set(sys.argv[1:]) or {"smoke", "spmm", "count", "graph", "model", "next"}
    t0 = time.time()
    done = []

    failed = []

    def run(name, fn, *a, **k):
        t = time.time()
        try:
            fn(*a, **k)
            torch.cuda.synchronize()
        except TypeError:
            raise
        except Exception as exc:  # noqa: BLE001 - keep going: the sanitizer's own report is what this run is for
            failed.append(name)
            print(f"[sanitize_run] {name}: PYTHON FAILURE {exc!r}", flush=True)
            return
        done.append(name)
        print(f"[sanitize_run] {name}: ok ({time.time() - t:.1f}s)", flush=True)

    if "smoke" in groups:
        import __graft_entry__ as entry
        run("smoke", entry.smoke)
    import tests.test_gpu_parity as tp
    if "spmm" in groups:
        run("spmm_vs_spec F=64 nv=3 long rows", tp.test_spmm_fanout_fanin_vs_spec, 64, 3, 96)
        run("spmm_vs_spec F=24 nv=1", tp.test_spmm_fanout_fanin_vs_spec, 24, 1, 0)
        run("spmm_vs_spec F=7 (scalar kernels)", tp.test_spmm_fanout_fanin_vs_spec, 7, 3, 0)
        run("spmm split operand F=128 w=32", tp.test_spmm_split_operand_and_column_chunks_bitwise, 128, 32, 3)
        run("spmm split operand F=24 w=8 nv=1", tp.test_spmm_split_operand_and_column_chunks_bitwise, 24, 8, 1)
    if "count" in groups:
        from protgram_directgcn_b200 import _native as nat
        for n in (1, 3, 4):
            for variant, code in tp.COUNT_VARIANTS.items():
                nat.load().pg_debug_count_variant(code)        # what the tests' `count_variant` fixture does
                try:
                    run(f"count n={n} {variant}", tp.test_count_and_extract_vs_c_oracle, n, variant)
                finally:
                    nat.load().pg_debug_count_variant(0)
        run("radix sort", tp.test_radix_sort_pairs_stable, 70_001, 40)
        run("unpack5", tp.test_unpack5_matches_host_packing, 4096)
    if "graph" in groups:
        import tests.test_partitioned_norm_gpu as tn
        import tempfile
        import pathlib
        with tempfile.TemporaryDirectory() as d:
            run("general edge table", tp.test_directed_ngram_graph_general_edge_table, pathlib.Path(d))
        for name in dir(tn):
            fn = getattr(tn, name)
            if name.startswith("test_row_block") and callable(fn) and not getattr(fn, "pytestmark", None):
                try:
                    run(name, fn)
                except TypeError:
                    pass     # parametrised / fixture-taking tests are covered by pytest proper
    if "model" in groups:
        run("layer gemms SIMT", tp.test_layer_gemms_vs_spec, 513, 16, 16, 0, 1)
        run("layer gemm fwd tcgen05", tp.test_layer_gemm_fwd_tensor_core_vs_spec, *tp.test_layer_gemm_fwd_tensor_core_vs_spec.pytestmark[0].args[1][0])
        run("layer gemm bwd tcgen05", tp.test_layer_gemm_bwd_tensor_core_vs_spec, *tp.test_layer_gemm_bwd_tensor_core_vs_spec.pytestmark[0].args[1][0])
        run("softmax nll", tp.test_softmax_nll_fused_vs_fp64, 300, 97, 100, True)
        run("l2 normalise", tp.test_l2_normalize_rows)
    if "next" in groups:
        import tests.test_next_rows as tx
        for name in dir(tx):
            fn = getattr(tx, name)
            if name.startswith("test_") and callable(fn):
                try:
                    run(name, fn)
                except TypeError:
                    pass
    print(f"[sanitize_run] {len(done)} cases ok, {len(failed)} failed {failed} in {time.time() - t0:.1f}s")
    sys.exit(1 if failed else 0)


if __name__ == "__main__":
    main()
