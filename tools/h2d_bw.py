"""H2D bandwidth of the step's input (67 MB pinned -> device), alone and while the GPU streams HBM at full rate on another
stream (what the e2e loop's prefetch sees), with one and with two copy streams."""
import torch
n = 67125248
x = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
big = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")   # 1 GB: add_ moves 2 GB per call (~0.35 ms)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def copy_ms(load, two, reps=10):
    for _ in range(2):
        d.copy_(x, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if load:
        for _ in range(reps * 8):
            big.add_(1)
    with torch.cuda.stream(s1):
        e0.record()
        for _ in range(reps):
            if two:
                h = n // 2
                d[:h].copy_(x[:h], non_blocking=True)
                with torch.cuda.stream(s2):
                    d[h:].copy_(x[h:], non_blocking=True)
            else:
                d.copy_(x, non_blocking=True)
        if two:
            s1.wait_stream(s2)
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for load in (False, True):
    for two in (False, True):
        ms = copy_ms(load, two)
        print(f"H2D 67 MB: load={load} two_streams={two}: {ms:.3f} ms = {n / ms / 1e6:.1f} GB/s")
