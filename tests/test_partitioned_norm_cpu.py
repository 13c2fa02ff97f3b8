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


def test_halo_plan_and_feature_chunks_pure_logic():
    """partitioned.halo_plan (which remote rows a block references, renumbered columns) and _feature_chunks are plain torch /
    Python: checked here without any kernel or process group."""
    import torch
    from protgram_directgcn_b200.host import partitioned as part
    n, world = 23, 3
    per = (n + world - 1) // world                                   # 8
    col = torch.tensor([0, 9, 8, 22, 9, 15, 8, 3], dtype=torch.int32)   # block of rank 1 owns rows [8, 16)
    need, counts, ext = part.halo_plan(col, 8, per, world)
    assert need.tolist() == [0, 3, 22] and counts.tolist() == [2, 0, 1]
    # own columns -> 0..per-1, halo columns -> per + position in `need`
    assert ext.tolist() == [8 + 0, 1, 0, 8 + 2, 1, 7, 0, 8 + 1] and ext.dtype == torch.int32
    # nothing remote: empty plan
    need0, counts0, ext0 = part.halo_plan(torch.tensor([8, 9, 15], dtype=torch.int32), 8, per, world)
    assert need0.numel() == 0 and counts0.tolist() == [0, 0, 0] and ext0.tolist() == [0, 1, 7]
    # feature chunks: multiples of 4 columns covering [0, f), one chunk unless pipelining is switched on and worth it
    assert part._feature_chunks(128, 1 << 21, 1) == [(0, 128)]
    saved = (part.PIPELINE_CHUNKS, part.PIPELINE_MIN_BYTES)
    try:
        part.PIPELINE_CHUNKS, part.PIPELINE_MIN_BYTES = 3, 0
        ch = part._feature_chunks(100, 10, 2)
        assert [c for c, _ in ch] == [0, 36, 72] and sum(w for _, w in ch) == 100 and all(w % 4 == 0 for _, w in ch)
        assert part._feature_chunks(6, 10, 2) == [(0, 6)]               # not a multiple of 4: never chunked
    finally:
        part.PIPELINE_CHUNKS, part.PIPELINE_MIN_BYTES = saved
    assert part.row_range(23, 2, 3) == (16, 23, 8)
