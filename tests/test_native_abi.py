"""The C-ABI library loads without a GPU and exports every symbol include/pgb200.h declares
(no compute calls here)."""
import os
import re

import pytest

from protgram_directgcn_b200 import _native as nat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "pgb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pg_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    lib = nat.load()
    names = _declared()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), f"{name} declared in pgb200.h but not exported by libpgb200.so"
    assert set(names) == set(nat.exported_symbols()), "ctypes signature table out of sync with the header"
    assert lib.pg_version() >= 100
    assert nat.query("pg_sort_pairs_ws_bytes", 10_000) > 0
    assert nat.query("pg_graph_extract_ws_bytes", 3, 21) > 21 ** 4 * 8
    tables = 148 * 224 * 1024                                                   # one packed table per CTA
    assert nat.query("pg_ngram_count_ws_bytes", 3, 21) >= 256 + 21 ** 4 * 8 + tables   # 8-bit lanes: + drain scratch
    assert nat.query("pg_ngram_count_ws_bytes", 1, 21) == 256 + tables          # strict lanes: status word + tables
    assert nat.query("pg_ngram_count_ws_bytes", 5, 21) == 256                   # 85.8 M bins: L2 REDs, no workspace
    # ... unless the caller sizes it for the partitioned variant: bucket entries (4 B / window) + scratch table + lists
    assert nat.query("pg_ngram_count_ws_bytes_for", 5, 21, 1 << 28) > 4 * (1 << 28) + 21 ** 6 * 8
    assert nat.query("pg_ngram_count_ws_bytes_for", 3, 21, 1 << 28) == nat.query("pg_ngram_count_ws_bytes", 3, 21)
    assert nat.query("pg_layer_gemm_bwd_data_tc_ws_bytes", 256, 256, 0) >= 3 * 16 * 2 * 4 * 256 * 16
    assert nat.query("pg_layer_gemm_bwd_weight_tc_ws_bytes", 168_000, 256, 256, 0) >= 771 * 256 * 4


def test_argument_errors_are_reported_not_crashed():
    lib = nat.load()
    # pure argument validation happens before any CUDA call
    rc = lib.pg_spmm_fanout(None, None, None, None, None, 2, 10, 8, None, 8, None, 24, 0, None, None)
    assert rc == -1 and b"nv must be 1 or 3" in lib.pg_last_error()
    rc = lib.pg_ngram_count(None, 0, 3, None, 21, None, None, None, 0, None)
    assert rc == -1
    with pytest.raises(nat.NativeError):
        nat.call("pg_sort_pairs", None, None, None, None, -5, 8, None, 0, None)
