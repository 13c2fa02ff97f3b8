#!/usr/bin/env python
"""Debug aid: weight-gradient tensor-core kernel, both shared-memory operand layouts, error statistics vs fp64."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from protgram_directgcn_b200 import _native as nat  # noqa: E402

DEV = torch.device("cuda", 0)


def run(n, f_in, f_out, has_res, mn):
    g = torch.Generator().manual_seed(1)
    z, x, dy = torch.randn(n, 3 * f_in, generator=g), torch.randn(n, f_in, generator=g), torch.randn(n, f_out, generator=g)
    gates = [torch.rand(n, generator=g) + 0.5 for _ in range(3)]
    k_data = 3 * f_in + (f_in if has_res else 0)
    k_ext = k_data + 3 + (1 if has_res else 0)
    cols = [z[:, v * f_in:(v + 1) * f_in] * gates[v][:, None] for v in range(3)]
    if has_res:
        cols.append(x)
    cols += [gates[0][:, None], gates[1][:, None], gates[2][:, None]]
    if has_res:
        cols.append(torch.ones(n, 1))
    a_ext = torch.cat(cols, 1).double()
    ref = a_ext.t() @ dy.double()
    d = lambda t: t.to(DEV).contiguous()
    zd, xd, dyd, gd = d(z), d(x), d(dy), [d(t) for t in gates]
    dw = torch.full((k_ext, f_out), float("nan"), device=DEV)
    need = nat.query("pg_layer_gemm_bwd_weight_tc_ws_bytes", n, f_in, f_out, has_res)
    ws = torch.zeros(need, dtype=torch.uint8, device=DEV)
    nat.load().pg_debug_tcw_layout(mn)
    st = nat.stream_ptr()
    nat.call("pg_layer_gemm_bwd_weight_tc", nat.ptr(zd), 3 * f_in, nat.ptr(xd), f_in, nat.ptr(gd[0]), nat.ptr(gd[1]), nat.ptr(gd[2]), 1,
             nat.ptr(dyd), f_out, n, f_in, f_out, has_res, nat.ptr(dw), nat.ptr(ws), ws.numel(), st)
    try:
        nat.call("pg_tc_check", nat.ptr(ws), need, st)
        flag = "ok"
    except Exception as exc:  # noqa: BLE001
        flag = f"WATCHDOG {exc}"
    out = dw.cpu().double()
    err = (out - ref).abs().max().item() / ref.abs().max().item()
    zeros = (out == 0).float().mean().item()
    # which rows / columns are right?
    row_ok = ((out - ref).abs().max(1).values / ref.abs().max() < 1e-4)
    col_ok = ((out - ref).abs().max(0).values / ref.abs().max() < 1e-4)
    print(f"n={n} f_in={f_in} f_out={f_out} res={has_res} mn_major={mn}: rel_err={err:.3e} zeros={zeros:.3f} nan={out.isnan().float().mean().item():.3f} "
          f"rows_ok={int(row_ok.sum())}/{k_ext} cols_ok={int(col_ok.sum())}/{f_out} {flag}", flush=True)
    if err > 1e-4 and n <= 64:
        print(" out[:4,:8]", out[:4, :8].tolist())
        print(" ref[:4,:8]", ref[:4, :8].tolist())


if __name__ == "__main__":
    for mn in (0, 1, 2, 3, 4):
        for shape in ((16, 32, 32, 0), (64, 32, 64, 1), (4096, 256, 256, 0)):
            run(*shape, mn)
    # timing at the C3 layer shape
    import time
    for mn in (0, 4):
        nat.load().pg_debug_tcw_layout(mn)
        n, f_in, f_out = 168_000, 256, 256
        z, x, dy = torch.randn(n, 3 * f_in, device=DEV), torch.randn(n, f_in, device=DEV), torch.randn(n, f_out, device=DEV)
        g = [torch.rand(n, device=DEV) + 0.5 for _ in range(3)]
        dw = torch.empty(3 * f_in + 3, f_out, device=DEV)
        need = nat.query("pg_layer_gemm_bwd_weight_tc_ws_bytes", n, f_in, f_out, 0)
        ws = torch.zeros(need, dtype=torch.uint8, device=DEV)
        st = nat.stream_ptr()
        fn = lambda: nat.call("pg_layer_gemm_bwd_weight_tc", nat.ptr(z), 3 * f_in, nat.ptr(x), f_in, nat.ptr(g[0]), nat.ptr(g[1]), nat.ptr(g[2]), 1,
                              nat.ptr(dy), f_out, n, f_in, f_out, 0, nat.ptr(dw), nat.ptr(ws), ws.numel(), st)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            fn()
        b.record()
        torch.cuda.synchronize()
        print(f"layout {mn}: C3 weight gradient {a.elapsed_time(b) / 10:.3f} ms", flush=True)
    nat.load().pg_debug_tcw_layout(0)
