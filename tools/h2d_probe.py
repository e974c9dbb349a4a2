"""Pinned host -> device copy bandwidth of this box (development aid): python tools/h2d_probe.py
One copy of S MiB, the same bytes as k back-to-back chunks on one stream, and as k chunks over two streams."""
import torch
torch.cuda.init()
def timed(fn, reps=8):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for mib in (1, 4, 16, 32, 64, 96, 256):
    n = mib << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
    ms = timed(lambda: d.copy_(h, non_blocking=True))
    print("H2D %4d MiB  1 copy: %.3f ms %.1f GB/s" % (mib, ms, n / ms / 1e6), flush=True)
n = 96 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
s2 = torch.cuda.Stream()
for k in (2, 3, 4, 8, 16):
    c = n // k
    def one():
        for i in range(k): d[i * c:(i + 1) * c].copy_(h[i * c:(i + 1) * c], non_blocking=True)
    ms = timed(one)
    print("H2D 96 MiB as %2d chunks, one stream: %.3f ms %.1f GB/s" % (k, ms, n / ms / 1e6), flush=True)
    def two():
        cur = torch.cuda.current_stream()
        s2.wait_stream(cur)
        for i in range(k):
            st = s2 if i & 1 else cur
            with torch.cuda.stream(st): d[i * c:(i + 1) * c].copy_(h[i * c:(i + 1) * c], non_blocking=True)
        cur.wait_stream(s2)
    ms = timed(two)
    print("H2D 96 MiB as %2d chunks, two streams: %.3f ms %.1f GB/s" % (k, ms, n / ms / 1e6), flush=True)
# fresh (never copied) pinned source each time: first-touch effects of the DMA mapping
srcs = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(4)]
torch.cuda.synchronize()
for i, hh in enumerate(srcs):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); d.copy_(hh, non_blocking=True); e1.record(); torch.cuda.synchronize()
    print("H2D 96 MiB first use of buffer %d: %.3f ms" % (i, e0.elapsed_time(e1)), flush=True)
