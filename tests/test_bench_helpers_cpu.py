"""CPU checks of bench.py's host-side helpers (no GPU work): the FASTA the `e2e_api` leg writes parses back to the bench corpus
through the package's own reader, both arms print the same `config` key set, the traffic files the roofline quotes are readable."""
import json
import os
import types

import numpy as np
import torch

import bench
from oracle import c_oracle, ngram_oracle


def test_bench_fasta_round_trips_through_the_reader(tmp_path, monkeypatch):
    import protgram_directgcn_b200 as pg
    monkeypatch.setattr(bench, "NSEQ", 64)
    seqs = ngram_oracle.synth_sequences(0, 64, bench.SEQ_LEN)
    for rank in (0, 1):        # rank 0's buffer starts with the leading space of global sequence #0
        buf = torch.from_numpy(c_oracle.pack_corpus(seqs, first_is_global_first=(rank == 0)))
        path = str(tmp_path / f"r{rank}.fasta")
        nbytes = bench.write_bench_fasta(types.SimpleNamespace(d_buf=buf, rank=rank), path)
        assert nbytes == os.path.getsize(path) == 64 * (17 + bench.SEQ_LEN + 1)
        parsed = list(pg.DataLoader.parse_sequences(path))
        assert [s for _, s in parsed] == seqs and parsed[0][0] == "P0000000" and parsed[-1][0] == "P0000063"


def test_both_arms_print_the_same_config_keys():
    ours = bench.step_config(nodes=1, unique_edges=2, pattern_nnz=3, l2_handling="x", cpu_affinity="y", multi_gpu="z", directgcn_step="g",
                             pipelining="p")
    ref = bench.step_config(nodes=1, unique_edges=2, note="n")
    assert set(ours) == set(ref) and ours["workload"] == ref["workload"] == bench.WORKLOAD
    try:
        bench.step_config(not_a_key=1)
    except AssertionError:
        pass
    else:
        raise AssertionError("unknown config keys must be refused")


def test_committed_traffic_captures_are_what_the_roofline_quotes():
    fo = bench.measured_traffic("spmm_fanout")
    cnt = bench.measured_traffic("ngram_count_smem_kernel", "r02_count_traffic.json")
    assert fo and cnt and fo["launches_per_call"] == 1
    # the SpMM moves far less than its algorithmic 37.9 GB through DRAM (hub rows live in L2); the count kernel reads the corpus once
    assert 10e9 < fo["bytes_per_call"] < 25e9 and 176e6 <= cnt["bytes_per_call"] < 200e6
    rec = json.load(open(os.path.join(bench.ROOT, "profiles", "r02_spmm_traffic.json")))
    assert all(l["kernel"].startswith("spmm_fan") for l in rec["launches"])
    assert bench.measured_traffic("no_such_kernel") is None and bench.measured_traffic("x", "missing.json") is None
