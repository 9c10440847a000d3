"""Time + check RoIAlign backward (TC path vs CUDA-core path) at the bench sizes."""
import os, sys, torch
sys.path.insert(0, ".")
from unsupervised_domain_adaptation_object_detection_implementation_b200 import functional as F_
from oracle import seeded
dev = "cuda"
N, C, H, W, R = 4, 2048, 64, 128, 2048
g = torch.Generator(device=dev).manual_seed(0)
feat = torch.relu(torch.randn(N, H, W, C, device=dev, generator=g)).to(torch.bfloat16).permute(0, 3, 1, 2).requires_grad_(True)
rois = seeded.synthetic_rois(R // N, N, H * 16, W * 16, 0).to(dev)
cot = torch.randn(R, C, 7, 7, device=dev, generator=g).to(torch.bfloat16)
def grad():
    out = F_.roi_align(feat, rois, 7, 1 / 16)
    (gi,) = torch.autograd.grad(out, feat, cot)
    return gi
def timeit(tag, n=10):
    out = F_.roi_align(feat, rois, 7, 1 / 16)
    for _ in range(3): torch.autograd.grad(out, feat, cot, retain_graph=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): torch.autograd.grad(out, feat, cot, retain_graph=True)
    e1.record(); torch.cuda.synchronize()
    print(f"{tag}: {e0.elapsed_time(e1)/n*1000:.0f} us per backward", flush=True)
g_tc = grad().float()
timeit("tc")
F_.set_option("roi_no_tc", 1)
g_cc = grad().float()
timeit("cuda-core")
err = (g_tc - g_cc).norm() / g_cc.norm()
print("rel fro err tc vs cuda-core:", float(err), "max abs", float((g_tc - g_cc).abs().max()), "ref max", float(g_cc.abs().max()))
F_.set_option("roi_no_tc", 0)
for d in (15, 32, 64):
    F_.set_option("roi_bwd_dbg", d)
    timeit(f"dbg={d}")
