"""Profiling driver: the FC1 contraction of the step (M=1024 RoIs, K=100352, N=1024) — forward, dgrad, wgrad."""
import sys, ctypes, torch
sys.path.insert(0, ".")
from unsupervised_domain_adaptation_object_detection_implementation_b200 import functional as F_, _lib
from unsupervised_domain_adaptation_object_detection_implementation_b200._lib import lib, check
dev = "cuda"
R, K, N = 1024, 100352, 1024
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(R, 1, 1, K, device=dev, generator=g).to(torch.bfloat16)
w = (torch.randn(N, K, device=dev, generator=g) * K ** -0.5).to(torch.bfloat16)
dz = torch.randn(R, 1, 1, N, device=dev, generator=g).to(torch.bfloat16)
y = torch.empty(R, 1, 1, N, device=dev, dtype=torch.bfloat16)
dx = torch.empty_like(x)
dw = torch.empty(N, 1, 1, K, device=dev, dtype=torch.float32)
desc = F_._conv_desc(R, 1, 1, K, N, 1, 1, 1, 0, "umma_bf16", torch.bfloat16, torch.bfloat16)
ws = F_.workspace(lib.da_conv_workspace_bytes(ctypes.byref(desc)), torch.device(dev), "conv")
P, S = F_._ptr, F_._stream
def fwd(): check(lib.da_conv_forward(ctypes.byref(desc), P(x), P(w), None, None, 1, 0.0, 0, P(y), P(ws), ws.numel(), S()))
def dgrad(): check(lib.da_conv_backward_data(ctypes.byref(desc), P(dz), P(w), 1.0, P(dx), P(ws), ws.numel(), S()))
def wgrad(): check(lib.da_conv_backward_weight(ctypes.byref(desc), P(x), P(dz), P(dw), P(ws), ws.numel(), S()))
flops = 2.0 * R * K * N
for name, fn in (("fwd", fwd), ("dgrad", dgrad), ("wgrad", wgrad)):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name}: {ms*1000:.0f} us  {flops/ms/1e9:.0f} TFLOP/s")
