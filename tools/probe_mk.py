"""Does a power-of-two row pitch of the A operand hurt the 1x1-conv GEMM (partition camping)?  Same GEMM at Cin = 2048 and
at neighbouring channel counts; time per FLOP should be flat if the address hash spreads the 128-byte row pieces."""
import sys, ctypes, torch
sys.path.insert(0, ".")
from unsupervised_domain_adaptation_object_detection_implementation_b200 import functional as F_
from unsupervised_domain_adaptation_object_detection_implementation_b200._lib import lib, check
dev = torch.device("cuda")
P, S = F_._ptr, F_._stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def run(M, Cin, Cout, n=12, relu=1, bias=False):
    x = torch.randn(M, Cin, device=dev).to(torch.bfloat16).view(M, 1, 1, Cin)
    w = (torch.randn(Cout, Cin, device=dev) * Cin ** -0.5).to(torch.bfloat16)
    y = torch.empty(M, 1, 1, Cout, device=dev, dtype=torch.bfloat16)
    desc = F_._conv_desc(M, 1, 1, Cin, Cout, 1, 1, 1, 0, "umma_bf16", torch.bfloat16, torch.bfloat16)
    ws = F_.workspace(lib.da_conv_workspace_bytes(ctypes.byref(desc)), dev, "conv")
    b_ = torch.randn(Cout, device=dev) if bias else None
    ts = []
    for i in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        check(lib.da_conv_forward(ctypes.byref(desc), P(x), P(w), None, P(b_), relu, 0.0, 0, P(y), P(ws), ws.numel(), S()))
        b.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(a.elapsed_time(b))
    ts.sort(); ms = ts[len(ts) // 2]
    fl = 2.0 * M * Cin * Cout
    print(f"relu={relu} bias={bias} M={M} Cin={Cin} Cout={Cout}: {ms * 1e3:7.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s")
for relu, bias in ((0, False), (1, False), (0, True), (1, True)):
    run(9472, 64, 512, relu=relu, bias=bias)
    run(9472, 2048, 512, relu=relu, bias=bias)
