"""A/B of the tensor-core RoIAlign forward at the step's size (roi_fwd_dbg bits: 1 no MMAs, 2 no output stores, 4 old 2-byte
staging stores); CUDA events, L2 flushed, median of 20, incl. the prep launches and the autograd wrapper."""
import sys, torch
sys.path.insert(0, ".")
from unsupervised_domain_adaptation_object_detection_implementation_b200 import functional as F_
from oracle import seeded
dev = "cuda"
N, C, H, W, R = 2, 2048, 64, 128, 1024
g = torch.Generator(device=dev).manual_seed(0)
feat = torch.relu(torch.randn(N, H, W, C, device=dev, generator=g)).to(torch.bfloat16).permute(0, 3, 1, 2)
rois = seeded.synthetic_rois(R // N, N, H * 16, W * 16, 0).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(tag, n=20):
    for _ in range(3): F_.roi_align(feat, rois, 7, 1 / 16)
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); F_.roi_align(feat, rois, 7, 1 / 16); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1000)
    ts.sort()
    print(f"{tag}: median {ts[n // 2]:.1f} us, min {ts[0]:.1f}", flush=True)
ref = None
for d in [int(a) for a in sys.argv[1:]] or [0, 4, 0, 4]:
    F_.set_option("roi_fwd_dbg", d)
    timeit(f"dbg={d}")
    out = F_.roi_align(feat, rois, 7, 1 / 16)
    if ref is None: ref = out.clone()
    elif d in (0, 4): print("   bit-identical to the first variant:", bool(torch.equal(out, ref)))
F_.set_option("roi_fwd_dbg", 0)
