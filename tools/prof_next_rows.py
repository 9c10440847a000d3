"""Driver for `ncu --set full` of the kernels of the SURVEY 8f rows built in round 2: two warm launches of the query-axis softmax
passes of the blocked NonLocalBlock on one [32768 x 4096] key block, and the RPN proposal stage of one image at the train
configuration.
    ncu --set full --clock-control none --import-source on -k regex:'colstats|col_apply|coldot|col_bwd|rpn_|nms_' \
        -o gpurun_out/r02_next python tools/prof_next_rows.py"""
import sys, torch
sys.path.insert(0, ".")
from unsupervised_domain_adaptation_object_detection_implementation_b200 import functional as F_
from unsupervised_domain_adaptation_object_detection_implementation_b200._lib import lib, check
from unsupervised_domain_adaptation_object_detection_implementation_b200.detection import AnchorGenerator
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
P, S = F_._ptr, F_._stream
Tq, Tk = 32768, 4096
s = torch.randn(Tq, Tk, device=dev, generator=g) * 3
p = torch.empty(Tq, Tk, dtype=torch.bfloat16, device=dev)
stats = torch.empty(2 * Tk, device=dev)
ws = torch.empty(lib.da_colsoftmax_workspace_bytes(Tq, Tk), dtype=torch.uint8, device=dev)
dp = torch.randn(Tq, Tk, device=dev, generator=g)
ds = torch.empty_like(p)
for _ in range(2):
    check(lib.da_colsoftmax_forward(P(s), Tq, Tk, Tk, P(p), 1, P(stats), 0, P(ws), ws.numel(), S()), "f")
    check(lib.da_colsoftmax_backward(P(p), 1, P(dp), Tq, Tk, Tk, P(ds), 1, P(ws), ws.numel(), S()), "b")
A, H, W = 15, 64, 128
cls = torch.randn(A, H, W, device=dev, generator=g) * 2
reg = torch.randn(4 * A, H, W, device=dev, generator=g) * 0.5
base = AnchorGenerator(strides=[16], ratios=[0.5, 1.0, 2.0], scales=[2, 4, 8, 16, 32]).base[0]
for _ in range(2):
    dets, count = F_.rpn_proposals(cls, reg, base, 16, (1024, 2048), 12000, 2000, 0.7, 0.0)
torch.cuda.synchronize()
print("proposals", int(count))
