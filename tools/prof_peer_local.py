"""The all-local update kernel of the peer optimizer (da_sgd_step_peer, DA_PEER_PUBLISH_BY_CALLER) on ONE device at the
slice size of an 8-GPU run of FC1 (12.8 M parameters, 8 gradient slots) - timing + driver for ncu."""
import sys, ctypes, torch
sys.path.insert(0, ".")
from unsupervised_domain_adaptation_object_detection_implementation_b200 import _lib, functional as F_, peer
from unsupervised_domain_adaptation_object_detection_implementation_b200._lib import lib, check
dev = "cuda"
world = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 8
n = 1024 * 100352
lo, hi, per = peer.slice_bounds(n, world, 0)
w = torch.randn(n, device=dev); mom = torch.zeros(per, device=dev)
grad = torch.randn(n, device=dev); staging = torch.randn(world * per, device=dev)
sh = torch.zeros(n, device=dev, dtype=torch.bfloat16)
flags = torch.zeros(_lib.DA_PEER_FLAG_INTS, dtype=torch.int32, device=dev)
state = torch.zeros(4, dtype=torch.int32, device=dev)
# one real rank + (world-1) staged slices that "arrived": mark their ready flags far in the future
flags[:world] = 1 << 30
state[3] = 1 << 30
gp = [grad.data_ptr() if q == 0 else staging.data_ptr() + 4 * (q * per - lo) for q in range(world)]
a = peer.make_args(w.data_ptr(), mom.data_ptr(), gp, [sh.data_ptr()] + [0] * (world - 1), None, [flags.data_ptr()] * world,
                   state.data_ptr(), n, world, 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for it in range(6):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(lib.da_sgd_step_peer(ctypes.byref(a), 1e-3, 0.9, 5e-4, 0, 0, _lib.DA_PEER_PUBLISH_BY_CALLER, None), "peer")
    e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ts = sorted(ts[1:]); t = ts[len(ts) // 2]
nb = (hi - lo) * (4 * world + 4 + 4 + 4 + 4 + 2)
print(f"peer update kernel, world {world}: slice {hi - lo} params, {t:.4f} ms, {nb / t / 1e6:.0f} GB/s of {nb / 1e6:.0f} MB "
      f"(reads {world} gradient slots + master + momentum, writes master + momentum + bf16); error flag {int(state[2])}")
