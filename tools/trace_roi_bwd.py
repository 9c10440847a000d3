"""Per-CTA timeline of the tensor-core RoIAlign backward (globaltimer stamps written by the kernel when
DA_ROI_BWD_TRACE holds the device address of a [ctas][8] u64 buffer)."""
import os, sys, torch
sys.path.insert(0, ".")
from unsupervised_domain_adaptation_object_detection_implementation_b200 import functional as F_
from oracle import seeded
dev = "cuda"
N, C, H, W, R = 2, int(os.environ.get('TRACE_C', 2048)), 64, 128, 1024
g = torch.Generator(device=dev).manual_seed(0)
feat = torch.relu(torch.randn(N, H, W, C, device=dev, generator=g)).to(torch.bfloat16).permute(0, 3, 1, 2).requires_grad_(True)
rois = seeded.synthetic_rois(R // N, N, H * 16, W * 16, 0).to(dev)
LAYOUT = os.environ.get("TRACE_LAYOUT", "rchw")      # rhwc: bin-major RoI tensor (TMA-fed MN-major gradient operand)
cot = torch.randn(R, C, 7, 7, device=dev, generator=g).to(torch.bfloat16)
if LAYOUT == "rhwc":
    cot = cot.permute(0, 2, 3, 1).contiguous()
out = F_.roi_align(feat, rois, 7, 1 / 16, out_layout=LAYOUT)
for _ in range(3): torch.autograd.grad(out, feat, cot, retain_graph=True)
ncta = 32 * (C // 256) * N
buf = torch.zeros(ncta + 64, 8, dtype=torch.int64, device=dev)
F_.set_option("roi_bwd_trace", buf.data_ptr())
if len(sys.argv) > 1: F_.set_option("roi_bwd_dbg", int(sys.argv[1]))
torch.autograd.grad(out, feat, cot, retain_graph=True)
torch.cuda.synchronize()
pt = buf[ncta:].cpu().double()
t = buf[:ncta].cpu().double()
t0 = t[:, 0].min()
start, comp, loop, end, pairs, smid = [(t[:, i] - (t0 if i < 4 else 0)) for i in range(6)]
print(f"kernel span {end.max()/1e3:.1f} us; CTAs {ncta}; pairs total {int(pairs.sum())}, max/CTA {int(pairs.max())}, mean {pairs.mean():.1f}")
print(f"compaction(first chunk) mean {(comp-start).mean()/1e3:.2f} us; pair phase mean {(loop-comp).mean()/1e3:.2f} us; epilogue mean {(end-loop).mean()/1e3:.2f} us; CTA mean {(end-start).mean()/1e3:.2f}")
pp = (loop - comp) / pairs.clamp(min=1)
print(f"per-pair (pair phase / pairs): mean {pp[pairs>8].mean():.0f} ns, p10 {pp[pairs>8].quantile(0.1):.0f}, p90 {pp[pairs>8].quantile(0.9):.0f}")
busy = torch.zeros(148)
for i in range(ncta): busy[int(smid[i]) % 148] += (end[i] - start[i])
print(f"SM busy: mean {busy.mean()/1e3:.1f} us, min {busy.min()/1e3:.1f}, max {busy.max()/1e3:.1f}; first CTA start spread {start.sort().values[min(147, ncta - 1)]/1e3:.1f} us")
last_start = start.max(); print(f"last CTA starts at {last_start/1e3:.1f} us; its duration {(end-start)[start.argmax()]/1e3:.1f}")

print(f"epilogue: tfull wait mean {(t[:,6]-t[:,2])[pairs>0].mean()/1e3:.2f} us; TMEM->global mean {(t[:,3]-t[:,6])[pairs>0].mean()/1e3:.2f} us")
base = pt[0, 0]
print("pair | prod:slot free | bld:raw here  ops free  built  synced | mma:ops ready  (gradient here)  committed   (SM cycles from first; builder columns: team 0 = even pairs)")
for i in range(24):
    r = pt[i] - base
    b = [f"{v:8.0f}" if v > -1e9 else "       -" for v in (r[3], r[4], r[5], r[6])]
    print(f"{i:3d} | {r[0]:8.0f} | {' '.join(b)} | {r[1]:8.0f} {r[7]:8.0f} {r[2]:8.0f}")
