"""Profiling driver: RoIAlign forward/backward at BASELINE config 4 (4 x 2048 x 64 x 128, 2048 RoIs)."""
import sys
import torch
sys.path.insert(0, ".")
from unsupervised_domain_adaptation_object_detection_implementation_b200 import functional as F_
from oracle import seeded

dt = torch.bfloat16 if (len(sys.argv) < 2 or sys.argv[1] == "bf16") else torch.float32
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = "cuda"
N, C, H, W, R = 4, 2048, 64, 128, 2048
g = torch.Generator(device=dev).manual_seed(0)
feat = torch.relu(torch.randn(N, H, W, C, device=dev, generator=g)).to(dt).permute(0, 3, 1, 2).requires_grad_(True)
rois = seeded.synthetic_rois(R // N, N, H * 16, W * 16, 0).to(dev)
cot = torch.randn(R, C, 7, 7, device=dev, generator=g).to(dt)
for _ in range(iters):
    out = F_.roi_align(feat, rois, 7, 1 / 16)
    (gin,) = torch.autograd.grad(out, feat, cot)
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record()
out = F_.roi_align(feat, rois, 7, 1 / 16)
e1.record()
(gin,) = torch.autograd.grad(out, feat, cot)
e2.record()
torch.cuda.synchronize()
print(f"dtype={dt} fwd {e0.elapsed_time(e1):.3f} ms  bwd {e1.elapsed_time(e2):.3f} ms")
