"""FC1 weight gradient: plain vs fused with the SGD step, each timed alone (CUDA events, L2 flushed) - also the driver for ncu."""
import sys, ctypes, torch
sys.path.insert(0, ".")
from unsupervised_domain_adaptation_object_detection_implementation_b200 import _lib, functional as F_
from unsupervised_domain_adaptation_object_detection_implementation_b200._lib import lib, check
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
P = F_._ptr
Rr, K, Nn = 1024, 100352, 1024
x = torch.randn(Rr, 1, 1, K, device=dev, generator=g).to(torch.bfloat16)
dz = torch.randn(Rr, 1, 1, Nn, device=dev, generator=g).to(torch.bfloat16)
dw = torch.empty(Nn * K, device=dev)
w = torch.randn(Nn * K, device=dev, generator=g)
buf = torch.zeros(Nn * K, device=dev)
sh = torch.zeros(Nn * K, device=dev, dtype=torch.bfloat16)
desc = F_._conv_desc(Rr, 1, 1, K, Nn, 1, 1, 1, 0, "umma_bf16", torch.bfloat16, torch.bfloat16)
ws = F_.workspace(lib.da_conv_workspace_bytes(ctypes.byref(desc)), torch.device(dev), "conv")
rec = _lib.SgdFuse(w.data_ptr(), buf.data_ptr(), sh.data_ptr(), 1e-3, 0.9, 5e-4, 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(fn, n=8):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return sum(ts[:4]) / 4
n_it = 1 if "--once" in sys.argv else 8
plain = t(lambda: check(lib.da_conv_backward_weight(ctypes.byref(desc), P(x), P(dz), P(dw), P(ws), ws.numel(), None)), n_it)
sgd = t(lambda: check(lib.da_sgd_step(P(w), P(dw), P(buf), Nn * K, 1e-3, 0.9, 5e-4, 0, P(sh), None)), n_it)
fused = t(lambda: check(lib.da_conv_backward_weight_sgd(ctypes.byref(desc), P(x), P(dz), ctypes.byref(rec), P(ws), ws.numel(), None)), n_it)
print(f"wgrad {plain:.4f} ms  sgd {sgd:.4f} ms  fused {fused:.4f} ms  ({Nn * K * 18 / fused / 1e6:.0f} GB/s of 18 B/param)")
rec1 = _lib.SgdFuse(w.data_ptr(), buf.data_ptr(), sh.data_ptr(), 1e-3, 0.9, 5e-4, 1)
f1 = t(lambda: check(lib.da_conv_backward_weight_sgd(ctypes.byref(desc), P(x), P(dz), ctypes.byref(rec1), P(ws), ws.numel(), None)), n_it)
rec2 = _lib.SgdFuse(w.data_ptr(), buf.data_ptr(), None, 1e-3, 0.9, 5e-4, 0)
f2 = t(lambda: check(lib.da_conv_backward_weight_sgd(ctypes.byref(desc), P(x), P(dz), ctypes.byref(rec2), P(ws), ws.numel(), None)), n_it)
print(f"fused first_step (no momentum read) {f1:.4f} ms; fused without bf16 copy {f2:.4f} ms")
