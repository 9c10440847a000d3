"""RoIAlign forward / backward alone at the bench size: CUDA events, L2 flushed between launches, median of 30."""
import os, sys, statistics, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unsupervised_domain_adaptation_object_detection_implementation_b200 import functional as F_
from oracle import seeded
dev = "cuda"
N, C, H, W, R = 2, 2048, 64, 128, 1024
g = torch.Generator(device=dev).manual_seed(0)
feat = torch.relu(torch.randn(N, H, W, C, device=dev, generator=g)).to(torch.bfloat16).permute(0, 3, 1, 2).requires_grad_(True)
rois = seeded.synthetic_rois(R // N, N, H * 16, W * 16, 0).to(dev)
cot = torch.randn(R, C, 7, 7, device=dev, generator=g).to(torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
tf, tb = [], []
for i in range(35):
    flush.fill_(i & 1)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e[0].record(); out = F_.roi_align(feat, rois, 7, 1 / 16); e[1].record()
    flush.fill_(i & 1)
    e[2].record(); torch.autograd.grad(out, feat, cot); e[3].record()
    torch.cuda.synchronize()
    if i >= 5:
        tf.append(e[0].elapsed_time(e[1])); tb.append(e[2].elapsed_time(e[3]))
print(f"roi_align fwd {statistics.median(tf):.4f} ms (min {min(tf):.4f})   bwd {statistics.median(tb):.4f} ms (min {min(tb):.4f})")
