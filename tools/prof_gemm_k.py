"""dgrad / wgrad time vs the contraction length: separates per-tile fixed cost from per-k-step cost."""
import sys, ctypes, torch
sys.path.insert(0, ".")
from unsupervised_domain_adaptation_object_detection_implementation_b200 import functional as F_
from unsupervised_domain_adaptation_object_detection_implementation_b200._lib import lib, check
dev = "cuda"
P, S = F_._ptr, F_._stream
def run(R, K, N):
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(R, 1, 1, K, device=dev, generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev, generator=g) * K ** -0.5).to(torch.bfloat16)
    dz = torch.randn(R, 1, 1, N, device=dev, generator=g).to(torch.bfloat16)
    dx = torch.empty_like(x); dw = torch.empty(N, 1, 1, K, device=dev, dtype=torch.float32)
    desc = F_._conv_desc(R, 1, 1, K, N, 1, 1, 1, 0, "umma_bf16", torch.bfloat16, torch.bfloat16)
    ws = F_.workspace(lib.da_conv_workspace_bytes(ctypes.byref(desc)), torch.device(dev), "conv")
    def dgrad(): check(lib.da_conv_backward_data(ctypes.byref(desc), P(dz), P(w), 1.0, P(dx), P(ws), ws.numel(), S()))
    def wgrad(): check(lib.da_conv_backward_weight(ctypes.byref(desc), P(x), P(dz), P(dw), P(ws), ws.numel(), S()))
    out = []
    for fn in (dgrad, wgrad):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / 5 * 1000)
    fl = 2.0 * R * K * N
    print(f"R={R} Cin={K} Cout={N}: dgrad {out[0]:.0f} us ({fl/out[0]/1e6:.0f} TF)  wgrad {out[1]:.0f} us ({fl/out[1]/1e6:.0f} TF)", flush=True)
# dgrad: K-loop over Cout; wgrad: K-loop over R
for N in (512, 1024): run(1024, 100352, N)

