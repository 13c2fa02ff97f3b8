import time, torch
dev = torch.device("cuda")
shapes = [((2, 168000), torch.int64), ((168000,), torch.float32), ((2, 168000), torch.int64), ((168000,), torch.float32),
          ((2, 344000), torch.int64), ((344000,), torch.float32), ((344000,), torch.float32), ((344000,), torch.float32)]
ts = [torch.zeros(s, dtype=d, device=dev) for s, d in shapes]
def a():
    return [t.cpu() for t in ts]
def b():
    host = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in ts]
    for h, t in zip(host, ts):
        h.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return host
keep = None
for name, fn in (("cpu()", a), ("pinned", b), ("cpu()", a), ("pinned", b)):
    torch.cuda.synchronize()
    for _ in range(3):
        keep = fn()
    t0 = time.perf_counter()
    for _ in range(20):
        keep = fn()
    torch.cuda.synchronize()
    print(name, (time.perf_counter() - t0) / 20 * 1e3, "ms")
