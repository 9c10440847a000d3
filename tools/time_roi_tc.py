import os, sys, torch
sys.path.insert(0, ".")
from unsupervised_domain_adaptation_object_detection_implementation_b200 import functional as F_
from oracle import seeded
dev = "cuda"
N, C, H, W, R = 4, 2048, 64, 128, 2048
g = torch.Generator(device=dev).manual_seed(0)
feat = torch.relu(torch.randn(N, H, W, C, device=dev, generator=g)).to(torch.bfloat16).permute(0, 3, 1, 2)
rois = seeded.synthetic_rois(R // N, N, H * 16, W * 16, 0).to(dev)
def run(tag, n=20):
    for _ in range(3): F_.roi_align(feat, rois, 7, 1 / 16)
    torch.cuda.synchronize()
    # GPU-side: batch many calls back-to-back so launch overhead overlaps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): F_.roi_align(feat, rois, 7, 1 / 16)
    e1.record(); torch.cuda.synchronize()
    print(f"{tag}: {e0.elapsed_time(e1)/n*1000:.0f} us per call")
for dbg in ("0", "1", "2", "4", "6", "7"):
    os.environ["DA_ROI_TC_DBG"] = dbg
    run("dbg=" + dbg)
small = rois.clone(); small[:, 3] = small[:, 1] + 64; small[:, 4] = small[:, 2] + 64
rois = small; os.environ["DA_ROI_TC_DBG"] = "0"; run("64px rois")
