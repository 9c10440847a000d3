"""One warm launch of the bf16 tensor-core RoIAlign forward and backward at the bench size (ncu target)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unsupervised_domain_adaptation_object_detection_implementation_b200 import functional as F_
from oracle import seeded
dev = "cuda"
N, C, H, W, R = 2, 2048, 64, 128, 1024
g = torch.Generator(device=dev).manual_seed(0)
feat = torch.relu(torch.randn(N, H, W, C, device=dev, generator=g)).to(torch.bfloat16).permute(0, 3, 1, 2).requires_grad_(True)
rois = seeded.synthetic_rois(R // N, N, H * 16, W * 16, 0).to(dev)
LAYOUT = os.environ.get("ROI_LAYOUT", "rhwc")      # memory order of the RoI tensor: rchw (reference) | rhwc (bin-major)
cot = torch.randn(R, C, 7, 7, device=dev, generator=g).to(torch.bfloat16)
if LAYOUT == "rhwc":
    cot = cot.permute(0, 2, 3, 1).contiguous()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    out = F_.roi_align(feat, rois, 7, 1 / 16, out_layout=LAYOUT)
    torch.autograd.grad(out, feat, cot)
torch.cuda.synchronize()
print("ok")
