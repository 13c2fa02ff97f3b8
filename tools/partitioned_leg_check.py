#!/usr/bin/env python
"""bench.py's `spmm_partitioned` leg on its own (torchrun, >= 2 GPUs): first checks, at 2^18 nodes per GPU, that the row
block obtained through the row-partitioned normalisation is bitwise the block cut from the replicated single-GPU
normalisation; then runs the leg at its bench size and prints its JSON.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/partitioned_leg_check.py"""
import json
import os
import sys
import types

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import bench
    from protgram_directgcn_b200 import _native as nat
    from protgram_directgcn_b200.host import graph_utils, partitioned as part
    pipe = types.SimpleNamespace(nat=nat, gu=graph_utils, dev=dev, rank=rank, world=world)
    log2 = 18 + (world - 1).bit_length()
    n, e, full = bench.rmat_graph(pipe, log2, 16)
    n2, e2, blk, ms = bench.rmat_row_block(pipe, dist, log2, 16)
    lo, hi, per = part.row_range(n, rank, world)
    ref = part.slice_rows(full["rowptr"], full["col"], [full["val_in"], full["val_out"], full["val_und"]], lo, hi, per)
    assert torch.equal(ref.rowptr, blk["rowptr"]) and torch.equal(ref.col, blk["col"])
    for a, k in zip(ref.vals, ("val_in", "val_out", "val_und")):
        assert torch.equal(a, blk[k]), k
    dist.barrier()
    if rank == 0:
        print(f"[ok] R-MAT 2^{log2} nodes over {world} GPUs: row block from the partitioned normalisation == block of the replicated one, "
              f"bitwise ({ms:.2f} ms)")
    del full, blk, ref
    torch.cuda.empty_cache()
    peak = bench.peaks()[0] if isinstance(bench.peaks(), tuple) else 6548.5
    leg = bench.spmm_partitioned_leg(pipe, peak, dist, int(os.environ.get("LEG_LOG2_PER_GPU", "21")))
    if rank == 0:
        print(json.dumps(leg))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
