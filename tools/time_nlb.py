"""Blocked NonLocalBlock attention at the C3 size of a 1024x2048 input (T = 32768 tokens, I = 256): event-timed forward and
forward+backward, and the column-softmax kernels alone against the HBM roofline.  python tools/time_nlb.py [T] [I] [block_k]"""
import json, sys, torch
sys.path.insert(0, ".")
from unsupervised_domain_adaptation_object_detection_implementation_b200 import functional as F_
T = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
I = int(sys.argv[2]) if len(sys.argv) > 2 else 256
BK = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
dev = "cuda"
g0 = torch.Generator(device=dev).manual_seed(0)
mk = lambda sc: (torch.randn(T, I, device=dev, generator=g0) * sc).to(torch.bfloat16).requires_grad_(True)
theta, phi, g = mk(0.2), mk(0.2), mk(1.0)
cot = torch.randn(T, I, device=dev, generator=g0).to(torch.bfloat16)
def ev(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]
def fwd():
    with torch.no_grad():
        return F_.nonlocal_attention_blocked(theta, phi, g, BK)
def fwdbwd():
    y = F_.nonlocal_attention_blocked(theta, phi, g, BK)
    torch.autograd.grad(y, (theta, phi, g), cot)
t_f, t_fb = ev(fwd), ev(fwdbwd)
gemm = 2.0 * T * T * I
out = {"T": T, "I": I, "block_k": BK, "fwd_ms": round(t_f, 3), "fwd_tflops": round(2 * gemm / t_f / 1e9, 1),
       "fwd_bwd_ms": round(t_fb, 3), "fwd_bwd_tflops": round(9 * gemm / t_fb / 1e9, 1),
       "note": "forward = 2 GEMMs of 2*T*T*I flop; forward+backward = 2 + 2 (recompute scores, dP) + 5 ... counted as 9 GEMM units"}
# the softmax kernels alone on one key block
s = torch.randn(T, BK, device=dev, generator=g0) * 3
p = torch.empty(T, BK, dtype=torch.bfloat16, device=dev)
stats = torch.empty(2 * BK, device=dev)
ws = torch.empty(F_.lib.da_colsoftmax_workspace_bytes(T, BK), dtype=torch.uint8, device=dev)
dp = torch.randn(T, BK, device=dev, generator=g0)
ds = torch.empty_like(p)
st = F_._stream
c = F_._code
t1 = ev(lambda: F_.check(F_.lib.da_colsoftmax_forward(F_._ptr(s), T, BK, BK, F_._ptr(p), c(p.dtype), F_._ptr(stats), 0, F_._ptr(ws), ws.numel(), st()), "f"), 20, 3)
t2 = ev(lambda: F_.check(F_.lib.da_colsoftmax_forward(F_._ptr(s), T, BK, BK, F_._ptr(p), c(p.dtype), F_._ptr(stats), 1, None, 0, st()), "f"), 20, 3)
t3 = ev(lambda: F_.check(F_.lib.da_colsoftmax_backward(F_._ptr(p), c(p.dtype), F_._ptr(dp), T, BK, BK, F_._ptr(ds), c(ds.dtype), F_._ptr(ws), ws.numel(), st()), "b"), 20, 3)
n = T * BK
out["colsoftmax_block"] = {"forward_ms": round(t1, 4), "forward_gbs": round(n * 10 / t1 / 1e6, 1), "apply_only_ms": round(t2, 4),
                           "apply_only_gbs": round(n * 6 / t2 / 1e6, 1), "backward_ms": round(t3, 4), "backward_gbs": round(n * 14 / t3 / 1e6, 1),
                           "bytes_per_element": "forward 4+4 read, 2 written; apply 4 read, 2 written; backward (2+4) x 2 read, 2 written"}
print(json.dumps(out))
