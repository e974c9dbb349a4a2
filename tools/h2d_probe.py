"""Pinned host -> device copy bandwidth of this box (development aid): python tools/h2d_probe.py"""
import torch, time
torch.cuda.init()
for mib in (32, 96, 256):
    h = torch.empty(mib << 20, dtype=torch.uint8).pin_memory()
    d = torch.empty(mib << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("H2D %d MiB pinned: %.3f ms  %.1f GB/s" % (mib, ms, (mib << 20) / ms / 1e6))
    e0.record()
    for _ in range(10): h.copy_(d, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("D2H %d MiB pinned: %.3f ms  %.1f GB/s" % (mib, ms, (mib << 20) / ms / 1e6))
import subprocess
print(subprocess.run("nvidia-smi topo -m | head -20; nvidia-smi -q | grep -i -A6 'GPU Link Info' | head -20; lscpu | grep -i -E 'numa|model name|^CPU\(s\)'", shell=True, capture_output=True, text=True).stdout)
