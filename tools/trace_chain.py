"""Per-group timeline of the instance-head chain kernel (CTA 0): GEMM phase, elementwise phase, grid barrier."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import unsupervised_domain_adaptation_object_detection_implementation_b200 as uda
from unsupervised_domain_adaptation_object_detection_implementation_b200 import functional as F_, da_heads
dev = torch.device("cuda")
R = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
torch.manual_seed(0)
m = da_heads.InstanceAlignmentHead().to(dev).train()
x = torch.relu(torch.randn(R, 1024, device=dev)).requires_grad_(True)
labels = (torch.arange(R, device=dev) >= R // 2).long()
for _ in range(3):
    loss, pred = m.forward_loss(x, labels)
    (loss + pred.sum() * 1e-3).backward()
torch.cuda.synchronize()
buf = torch.zeros(256, dtype=torch.int64, device=dev)
names_f = ["proj", "S", "softmax", "P*g", "mask+res", "fc1", "fc2", "fc3", "CE"]
names_b = ["ce_bwd", "fc3 bwd", "fc2 bwd", "fc1 bwd", "mask bwd", "dP,dg", "softmax bwd", "dtheta,dphi", "dWproj,dx"]
for which, names in (("forward", names_f), ("backward", names_b)):
    loss, pred = m.forward_loss(x, labels)
    torch.cuda.synchronize()
    buf.zero_()
    F_.set_option("chain_trace", buf.data_ptr())
    if which == "forward":
        loss, pred = m.forward_loss(x, labels)
    else:
        (loss + pred.sum() * 1e-3).backward()
    torch.cuda.synchronize()
    F_.set_option("chain_trace", 0)
    t = buf.cpu().tolist()
    t0 = t[0]
    print(f"{which}: total {(max(t[:128]) - t0) / 1e3:.1f} us")
    for g, n in enumerate(names):
        a, gemm, ew, bar = t[3 * g], t[3 * g + 1], t[3 * g + 2], t[3 * g + 3]
        if gemm == 0:
            break
        roles = [(t[128 + 4 * g + k] - a) / 1e3 for k in range(4)]
        print(f"  {n:12s} gemm {(gemm - a) / 1e3:6.2f}  elementwise {(ew - gemm) / 1e3:6.2f}  barrier {(bar - ew) / 1e3:6.2f} us"
              f"   role done at: tma {roles[0]:5.2f} mma {roles[1]:5.2f} epi-first {roles[2]:5.2f} epi-last {roles[3]:5.2f}")
