"""How much of a read-only working set does the B200 L2 keep from one kernel to the next?  Repeated full reads (torch.sum, plain
LDG) of buffers of 16..256 MB: effective GB/s per pass (above the HBM peak = served from L2)."""
import torch
for mb in (16, 32, 48, 56, 64, 72, 96, 128, 256):
    x = torch.ones(mb * 1024 * 1024 // 2, dtype=torch.bfloat16, device="cuda")
    for _ in range(3):
        x.sum()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 30
    e0.record()
    for _ in range(n):
        x.sum()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print("%4d MB: %.1f us per pass, %.0f GB/s" % (mb, ms * 1e3, mb * 1.048576 / ms))
