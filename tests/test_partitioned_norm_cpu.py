"""Host logic of the row-partitioned normalisation on the kernels' executable spec: random small graphs (isolated nodes, self
loops only, more ranks than nodes, no edges at all) played rank by rank against the full-graph oracle."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from protgram_directgcn_b200 import _native as nat
from tests import kernel_spec
from tests.test_multirank_gloo import check_blocks_against_oracle


@settings(max_examples=120, deadline=None)
@given(st.integers(min_value=1, max_value=12), st.integers(min_value=1, max_value=6), st.integers(min_value=0, max_value=2 ** 31 - 1),
       st.sampled_from([0.0, 0.1, 0.4, 1.0]))
def test_played_ranks_match_oracle_on_random_small_graphs(n, world, seed, density):
    import pytest
    from protgram_directgcn_b200.host.partitioned import row_range
    import tests.test_partitioned_norm_gpu as gpu_tests
    mp = pytest.MonkeyPatch()
    try:
        kernel_spec.install(mp, nat)
        mp.setattr(gpu_tests, "DEV", "cpu")
        rng = np.random.default_rng(seed)
        mask = rng.random((n, n)) < density
        src, dst = np.nonzero(mask)
        src, dst = src.astype(np.int64), dst.astype(np.int64)
        cnt = rng.integers(1, 9, size=src.size).astype(np.int64)
        blocks, _ = gpu_tests._play_ranks(src, dst, cnt.astype(np.float32), n, world, seed=seed % 1000)
        host = {}
        for r, b in enumerate(blocks):
            lo, hi, per = row_range(n, r, world)
            rp = np.full(per + 1, b["pattern_nnz"], dtype=np.int64)
            rp[: hi - lo + 1] = b["rowptr"].numpy()
            host[r] = {k: v.numpy() for k, v in b.items() if torch.is_tensor(v)}
            host[r]["rowptr"] = rp
        if src.size:
            check_blocks_against_oracle(host, src, dst, cnt, n, world)
        else:   # no edges: the oracle returns empty matrices (reference :205-207); the blocks hold the identity pattern
            assert all(np.array_equal(np.diff(host[r]["rowptr"][: row_range(n, r, world)[1] - row_range(n, r, world)[0] + 1]),
                                      np.ones(row_range(n, r, world)[1] - row_range(n, r, world)[0], dtype=np.int64)) for r in host)
    finally:
        mp.undo()
