"""Round-2 driver for `ncu --set full`: ONE warm launch of every hot kernel of the step at the step's sizes
(RoIAlign fwd/bwd, FC1 fwd / dgrad / fused wgrad+SGD, instance-head chain fwd/bwd with its feeding layer).
    ncu --set full --clock-control none --import-source on -k regex:'roi_align_(fwd|bwd)_tc|umma_nt_kernel|umma_tn_kernel|chain_kernel' \
        -o gpurun_out/r02_hot python tools/prof_r02.py"""
import sys, ctypes, torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from unsupervised_domain_adaptation_object_detection_implementation_b200 import _lib, functional as F_, da_heads, hotpath
from unsupervised_domain_adaptation_object_detection_implementation_b200._lib import lib, check
from oracle import seeded
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
P, S = F_._ptr, F_._stream
N, C, H, W, R = 2, 2048, 64, 128, 1024
feat = torch.relu(torch.randn(N, H, W, C, device=dev, generator=g)).to(torch.bfloat16).permute(0, 3, 1, 2).requires_grad_(True)
rois = seeded.synthetic_rois(R // N, N, H * 16, W * 16, 0).to(dev)
for _ in range(2):
    out = F_.roi_align(feat, rois, 7, 1 / 16)
    cot = torch.randn(out.shape, device=dev, generator=g).to(torch.bfloat16)
    torch.autograd.grad(out, feat, cot)
del out, cot, feat
Rr, K, Nn = 1024, 100352, 1024
x = torch.randn(Rr, 1, 1, K, device=dev, generator=g).to(torch.bfloat16)
w = (torch.randn(Nn, K, device=dev, generator=g) * K ** -0.5)
wb = w.to(torch.bfloat16)
dz = torch.randn(Rr, 1, 1, Nn, device=dev, generator=g).to(torch.bfloat16)
y = torch.empty(Rr, 1, 1, Nn, device=dev, dtype=torch.bfloat16)
dx = torch.empty_like(x)
desc = F_._conv_desc(Rr, 1, 1, K, Nn, 1, 1, 1, 0, "umma_bf16", torch.bfloat16, torch.bfloat16)
ws = F_.workspace(lib.da_conv_workspace_bytes(ctypes.byref(desc)), torch.device(dev), "conv")
buf = torch.zeros(Nn * K, device=dev)
rec = _lib.SgdFuse(w.data_ptr(), buf.data_ptr(), wb.data_ptr(), 1e-3, 0.9, 5e-4, 0)
for _ in range(2):
    check(lib.da_conv_forward(ctypes.byref(desc), P(x), P(wb), None, None, 1, 0.0, 0, P(y), P(ws), ws.numel(), S()))
    check(lib.da_conv_backward_data(ctypes.byref(desc), P(dz), P(wb), 1.0, P(dx), P(ws), ws.numel(), S()))
    check(lib.da_conv_backward_weight_sgd(ctypes.byref(desc), P(x), P(dz), ctypes.byref(rec), P(ws), ws.numel(), S()))
del x, dx, y, dz, w, wb, buf
torch.manual_seed(0)
fcs = hotpath.SharedFCs(64, 2, 1024).to(dev)
head = da_heads.InstanceAlignmentHead().to(dev).train()
roi = torch.relu(torch.randn(R, 64, 2, 2, device=dev)).to(torch.bfloat16).requires_grad_(True)
labels = (torch.arange(R, device=dev) >= R // 2).int()
for _ in range(2):
    xin, pre = fcs.split(roi)
    loss, pred = head.forward_loss(None, labels, pre=(xin,) + pre)
    (loss + pred.sum() * 1e-3).backward()
torch.cuda.synchronize()
print("ok")
