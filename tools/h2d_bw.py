import torch, time
x = torch.empty(67125248, dtype=torch.uint8).pin_memory()
d = torch.empty_like(x, device="cuda")
s = torch.cuda.Stream()
for n in (1, 2):
    with torch.cuda.stream(s):
        for _ in range(3): d.copy_(x, non_blocking=True)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): d.copy_(x, non_blocking=True)
        e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"H2D 67 MB pinned: {ms:.3f} ms = {67.125248 / ms:.1f} GB/s")
