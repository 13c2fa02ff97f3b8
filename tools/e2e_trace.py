#!/usr/bin/env python
"""Diagnostic: per-step wall time of bench.py's e2e step with a host-side phase split (sync after each phase)."""
import os, sys, time, gc
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench


def main():
    torch.cuda.set_device(0)
    pipe = bench.B200Pipeline(0, 1, torch.device("cuda", 0))
    for _ in range(3):
        pipe.step_resident()
    torch.cuda.synchronize()
    sync = "--sync" in sys.argv
    nsteps = 400 if '--long' in sys.argv else 24
    for i in range(nsteps):
        if i == 12:
            torch.cuda.synchronize()
            time.sleep(0.5)          # like the barrier between warm-up and timed region
        t0 = time.perf_counter()
        if pipe.h_buf is None:
            pipe.h_buf = pipe.d_buf.cpu().pin_memory()
            pipe.up = pipe.corpus.CorpusUploader(pipe.dev)
            pipe.up.submit(pipe.h_buf)

        def prefetch_next():
            pipe.up.release()
            pipe.up.submit(pipe.h_buf)
        d_buf = pipe.up.acquire()
        if sync: torch.cuda.synchronize()
        t1 = time.perf_counter()
        graph = pipe.build(d_buf, materialise_host=True, after_count=prefetch_next)
        if sync: torch.cuda.synchronize()
        t2 = time.perf_counter()
        loss, emb = pipe.train_and_extract(graph)
        graph.node_sequences
        if sync: torch.cuda.synchronize()
        t3 = time.perf_counter()
        emb_host = emb.cpu().numpy()
        l = float(loss.item())
        t4 = time.perf_counter()
        if nsteps > 24 and (t4 - t0) < 4.5e-3 and i > 3:
            continue
        print(f"step {i:2d}: total {1e3*(t4-t0):7.3f} ms | acquire {1e3*(t1-t0):6.3f} build {1e3*(t2-t1):6.3f} train {1e3*(t3-t2):6.3f} readback {1e3*(t4-t3):6.3f} | gc {gc.get_count()}")


if __name__ == "__main__":
    main()
