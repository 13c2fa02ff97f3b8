#!/usr/bin/env python
"""Time the layer forward transform: SIMT fp32 (gemm.cu) vs tcgen05 3xTF32 (gemm_tc.cu) at C3/C5-like shapes."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from protgram_directgcn_b200 import _native as nat

def run(n, f_in, f_out, has_res, iters=10):
    dev = "cuda"
    z, x = torch.randn(n, 3 * f_in, device=dev), torch.randn(n, f_in, device=dev)
    g = [torch.rand(n, device=dev) + 0.5 for _ in range(3)]
    k_ext = 3 * f_in + (f_in if has_res else 0) + 3 + (1 if has_res else 0)
    w = torch.randn(k_ext, f_out, device=dev) * 0.1
    c = torch.randn(n, f_out, device=dev)
    h = torch.empty(n, f_out, device=dev)
    ws = torch.empty(nat.query("pg_layer_gemm_fwd_tc_ws_bytes", f_in, f_out, has_res), dtype=torch.uint8, device=dev)
    st = nat.stream_ptr()
    ai = int((not has_res) and f_in == f_out)
    simt = lambda: nat.call("pg_layer_gemm_fwd", nat.ptr(z), 3 * f_in, nat.ptr(x), f_in, nat.ptr(g[0]), nat.ptr(g[1]), nat.ptr(g[2]), 1,
                            nat.ptr(w), nat.ptr(c), f_out, n, f_in, f_out, has_res, ai, 0.01, nat.ptr(h), f_out, st)
    tc = lambda: nat.call("pg_layer_gemm_fwd_tc", nat.ptr(z), 3 * f_in, nat.ptr(x), f_in, nat.ptr(g[0]), nat.ptr(g[1]), nat.ptr(g[2]), 1,
                          nat.ptr(w), nat.ptr(c), f_out, n, f_in, f_out, has_res, ai, 0.01, nat.ptr(h), f_out, nat.ptr(ws), ws.numel(), st)
    out = {}
    for name, fn in (("simt", simt), ("tc", tc)):
        for _ in range(3):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / iters
        out[name] = (round(ms, 4), round(2.0 * n * k_ext * f_out / ms / 1e9, 1))
    nat.call("pg_layer_gemm_fwd_tc_check", nat.ptr(ws), f_in, f_out, has_res, st)
    print(f"N={n} F_in={f_in} F_out={f_out} K_ext={k_ext}: SIMT {out['simt'][0]} ms ({out['simt'][1]} TFLOP/s fp32)  "
          f"tcgen05 3xTF32 {out['tc'][0]} ms ({out['tc'][1]} TFLOP/s effective)")

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "c2"):
        run(8401, 64, 256, 1); run(8401, 256, 128, 1); run(8401, 128, 64, 1)
    if which in ("all", "c3"):
        run(160000, 256, 256, 0)
    if which in ("all", "c5"):
        run(2 ** 21, 128, 128, 0)
