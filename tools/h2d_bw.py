"""Raw pinned H2D / D2H bandwidth of the box (context for bench.py's e2e number)."""
import torch, time
dev = torch.device("cuda:0")
for mb in (16, 113, 512):
    h = torch.empty(mb * 1024 * 1024, dtype=torch.uint8).pin_memory()
    d = torch.empty_like(h, device=dev)
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    e0.record()
    for _ in range(10):
        h.copy_(d, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / 10
    print(f"{mb} MiB: H2D {mb * 1.048576 / ms:.1f} GB/s   D2H {mb * 1.048576 / ms2:.1f} GB/s")
