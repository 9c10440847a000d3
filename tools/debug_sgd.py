import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unsupervised_domain_adaptation_object_detection_implementation_b200 import _lib, functional as F_
from unsupervised_domain_adaptation_object_detection_implementation_b200._lib import check, lib
DEV = "cuda"
N, Cin, Cout = 64, int(sys.argv[1]) if len(sys.argv) > 1 else 76800, 256
g = torch.Generator(device=DEV).manual_seed(1)
n = Cout * Cin
w_a = torch.randn(n, device=DEV, generator=g); w_b = w_a.clone()
buf_a, buf_b = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
sh_a = torch.zeros(n, device=DEV, dtype=torch.bfloat16); sh_b = torch.zeros(n, device=DEV, dtype=torch.bfloat16)
dw = torch.empty(n, device=DEV)
desc = F_._conv_desc(N, 1, 1, Cin, Cout, 1, 1, 1, 0, "umma_bf16", torch.bfloat16, torch.bfloat16)
ws = F_.workspace(lib.da_conv_workspace_bytes(ctypes.byref(desc)), torch.device(DEV), "conv")
x = torch.randn(N, 1, 1, Cin, device=DEV, generator=g).to(torch.bfloat16)
dz = torch.randn(N, 1, 1, Cout, device=DEV, generator=g).to(torch.bfloat16)
check(lib.da_conv_backward_weight(ctypes.byref(desc), F_._ptr(x), F_._ptr(dz), F_._ptr(dw), F_._ptr(ws), ws.numel(), None), "wgrad")
check(lib.da_sgd_step(F_._ptr(w_a), F_._ptr(dw), F_._ptr(buf_a), n, 0.05, 0.9, 5e-4, 1, F_._ptr(sh_a), None), "sgd")
rec = _lib.SgdFuse(w_b.data_ptr(), buf_b.data_ptr(), sh_b.data_ptr(), 0.05, 0.9, 5e-4, 1)
check(lib.da_conv_backward_weight_sgd(ctypes.byref(desc), F_._ptr(x), F_._ptr(dz), ctypes.byref(rec), F_._ptr(ws), ws.numel(), None), "wgrad_sgd")
torch.cuda.synchronize()
bad = (w_a != w_b).view(Cout, Cin)
print("mismatching elements:", int(bad.sum()), "of", n)
cols = bad.any(0).view(-1, 32).any(1)          # per 32-col slab
tiles = cols.view(-1, 8)                        # [ci tile][chunk]
bt = tiles.any(1).nonzero().flatten().tolist()
print("bad ci tiles:", bt[:40], "... total", len(bt), "of", tiles.shape[0])
print("bad chunks histogram:", tiles.sum(0).tolist())
rows = bad.any(1).view(2, 128).sum(1).tolist()
print("bad rows per co tile:", rows)
if bt:
    t = bt[0]
    sub = bad[:, t * 256:(t + 1) * 256]
    print("first bad tile", t, "bad per chunk", sub.view(Cout, 8, 32).any(2).sum(0).tolist(), "bad rows", int(sub.any(1).sum()))
    i = sub.nonzero()[0].tolist()
    print("example", i, float(w_a.view(Cout, Cin)[i[0], t * 256 + i[1]]), float(w_b.view(Cout, Cin)[i[0], t * 256 + i[1]]),
          "buf", float(buf_a.view(Cout, Cin)[i[0], t * 256 + i[1]]), float(buf_b.view(Cout, Cin)[i[0], t * 256 + i[1]]))
