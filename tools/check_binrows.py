"""Reads the RoIAlign workspace after a forward at the bench size and prints the histogram of k-steps per (RoI, tile) pair that the
bin-major tensor-core backward derives from the Wy pad column (roi_prep_kernel)."""
import sys, torch
sys.path.insert(0, ".")
from unsupervised_domain_adaptation_object_detection_implementation_b200 import functional as F_
from unsupervised_domain_adaptation_object_detection_implementation_b200._lib import lib
from oracle import seeded
dev = "cuda"
N, C, H, W, R = 2, 256, 64, 128, 1024
feat = torch.relu(torch.randn(N, H, W, C, device=dev)).to(torch.bfloat16).permute(0, 3, 1, 2)
rois = seeded.synthetic_rois(R // N, N, H * 16, W * 16, 0).to(dev)
out = F_.roi_align(feat, rois, 7, 1 / 16, out_layout="rhwc")
torch.cuda.synchronize()
ws = F_.workspace(lib.da_roi_align_workspace_bytes(R, H, W), feat.device, "roi").cpu()
meta = ws[16:16 + R * 32].view(torch.int32).view(R, 8)
toff = 16 + (R * 32 + 255) // 256 * 256
tab = ws[toff:toff + R * (H + W) * 8 * 4].view(torch.int32).view(R, H + W, 8)
hist, pairs, bins = {1: 0, 2: 0, 3: 0, 4: 0}, 0, 0
for r in range(R):
    b, gh, gw, y_lo, ny, x_lo, nx, cnt = meta[r].tolist()
    if ny <= 0 or nx <= 0:
        continue
    for ty in range(y_lo // 16, (y_lo + ny - 1) // 16 + 1):
        ya, yb = max(y_lo, ty * 16), min(y_lo + ny, ty * 16 + 16)
        fa = int(tab[r, ya - y_lo, 7]) & 255
        lb = (int(tab[r, yb - 1 - y_lo, 7]) >> 8) & 255
        pa = min(fa, 6); pb = max(min(lb, 7), pa + 1)
        ks = ((pb - pa) * 7 + 15) >> 4
        ntx = (x_lo + nx - 1) // 16 - x_lo // 16 + 1
        hist[ks] += ntx; pairs += ntx; bins += ntx * min(16 * ks, 49 - pa * 7)
print("k-step histogram", hist, "pairs", pairs, "mean k-steps", sum(k * v for k, v in hist.items()) / pairs, "mean bins fetched", bins / pairs)
print("sample pads", [hex(int(v)) for v in tab[0, :8, 7]], meta[0].tolist())
