"""Ablations of the tensor-core RoIAlign backward at the step's size (roi_bwd_dbg bits; results are wrong when set)."""
import sys, torch
sys.path.insert(0, ".")
from unsupervised_domain_adaptation_object_detection_implementation_b200 import functional as F_
from oracle import seeded
dev = "cuda"
N, C, H, W, R = 2, 2048, 64, 128, 1024
g = torch.Generator(device=dev).manual_seed(0)
feat = torch.relu(torch.randn(N, H, W, C, device=dev, generator=g)).to(torch.bfloat16).permute(0, 3, 1, 2).requires_grad_(True)
rois = seeded.synthetic_rois(R // N, N, H * 16, W * 16, 0).to(dev)
cot = torch.randn(R, C, 7, 7, device=dev, generator=g).to(torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = F_.roi_align(feat, rois, 7, 1 / 16)
def timeit(tag, n=20):
    for _ in range(3): torch.autograd.grad(out, feat, cot, retain_graph=True)
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.autograd.grad(out, feat, cot, retain_graph=True); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1000)
    ts.sort()
    print(f"{tag}: median {ts[len(ts)//2]:.1f} us, min {ts[0]:.1f}", flush=True)
ref = torch.autograd.grad(out, feat, cot, retain_graph=True)[0].float()
for d in [int(a) for a in sys.argv[1:]] or [0, 32, 64, 16, 31, 15, 2, 4, 1]:
    F_.set_option("roi_bwd_dbg", d)
    timeit(f"dbg={d}")
    if d == 0:
        gi = torch.autograd.grad(out, feat, cot, retain_graph=True)[0].float()
        print("   max abs diff vs shipped:", float((gi - ref).abs().max()))
F_.set_option("roi_bwd_dbg", 0)
