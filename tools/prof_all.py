"""One launch of every hot kernel of the step at the step's sizes (driver for `ncu --set full`)."""
import sys, ctypes, torch
sys.path.insert(0, ".")
from unsupervised_domain_adaptation_object_detection_implementation_b200 import functional as F_, optim
from unsupervised_domain_adaptation_object_detection_implementation_b200._lib import lib, check
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
P, S = F_._ptr, F_._stream
# --- RoIAlign forward / backward: 2 images, 1024 RoIs, C5 = [2, 2048, 64, 128] bf16
N, C, H, W, R = 2, 2048, 64, 128, 1024
feat = torch.relu(torch.randn(N, H, W, C, device=dev, generator=g)).to(torch.bfloat16).permute(0, 3, 1, 2).requires_grad_(True)
cg = torch.Generator().manual_seed(1)
u = torch.rand(R, 4, generator=cg)
x1, y1 = u[:, 0] * (W * 16 - 33), u[:, 1] * (H * 16 - 33)
wh = torch.exp(torch.log(torch.tensor(16.0)) + u[:, 2:] * (torch.log(torch.tensor(512.0)) - torch.log(torch.tensor(16.0))))
rois = torch.stack([(torch.arange(R) // (R // N)).float(), x1, y1, torch.clamp(x1 + wh[:, 0], max=W * 16),
                    torch.clamp(y1 + wh[:, 1], max=H * 16)], 1).to(dev)
out = F_.roi_align(feat, rois, 7, 1 / 16)
cot = torch.randn(out.shape, device=dev, generator=g).to(torch.bfloat16)
torch.autograd.grad(out, feat, cot)
del out, cot, feat
# --- FC1: [1024, 100352] x [100352 -> 1024]
Rr, K, Nn = 1024, 100352, 1024
x = torch.randn(Rr, 1, 1, K, device=dev, generator=g).to(torch.bfloat16)
w = (torch.randn(Nn, K, device=dev, generator=g) * K ** -0.5).to(torch.bfloat16)
dz = torch.randn(Rr, 1, 1, Nn, device=dev, generator=g).to(torch.bfloat16)
y = torch.empty(Rr, 1, 1, Nn, device=dev, dtype=torch.bfloat16)
dx = torch.empty_like(x)
dw = torch.empty(Nn, 1, 1, K, device=dev, dtype=torch.float32)
desc = F_._conv_desc(Rr, 1, 1, K, Nn, 1, 1, 1, 0, "umma_bf16", torch.bfloat16, torch.bfloat16)
ws = F_.workspace(lib.da_conv_workspace_bytes(ctypes.byref(desc)), torch.device(dev), "conv")
check(lib.da_conv_forward(ctypes.byref(desc), P(x), P(w), None, None, 1, 0.0, 0, P(y), P(ws), ws.numel(), S()))
check(lib.da_conv_backward_data(ctypes.byref(desc), P(dz), P(w), 1.0, P(dx), P(ws), ws.numel(), S()))
check(lib.da_conv_backward_weight(ctypes.byref(desc), P(x), P(dz), P(dw), P(ws), ws.numel(), S()))
del x, dx, y, dz
# --- fused SGD over the FC1 weight
p = torch.nn.Parameter(torch.randn(Nn, K, device=dev, generator=g))
p.grad = dw.view(Nn, K)
opt = optim.FusedSGD([p], lr=1e-3, momentum=0.9, weight_decay=5e-4)
opt.step(); opt.step()
torch.cuda.synchronize()
print("ok")
