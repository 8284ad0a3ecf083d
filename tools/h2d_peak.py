"""pinned host -> device copy bandwidth of this box (what bounds e2e): one stream, large and chunk-sized copies"""
import torch, time
for mb in (48, 256, 1024, 3072):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True); h.fill_(1)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(2): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    reps = max(2, 4096 // mb)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print(f"H2D {mb} MiB x{reps}: {n * reps / (e0.elapsed_time(e1) / 1e3) / 1e9:.1f} GB/s")
