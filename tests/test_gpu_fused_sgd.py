"""Weight gradient fused with the SGD step (da_conv_backward_weight_sgd, the epilogue of the tcgen05 weight-gradient
kernel) against the two-kernel path it replaces: da_conv_backward_weight -> da_sgd_step.  Same gradient bits, same
operation order of the update => master, momentum and bf16 operand copy must be bit-identical."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu

import unsupervised_domain_adaptation_object_detection_implementation_b200 as uda  # noqa: E402
from unsupervised_domain_adaptation_object_detection_implementation_b200 import _lib, functional as F_, optim  # noqa: E402
from unsupervised_domain_adaptation_object_detection_implementation_b200._lib import check, lib  # noqa: E402

DEV = "cuda"
LR, MU, WD = 0.05, 0.9, 5e-4


@pytest.mark.parametrize("N,H,W,Cin,Cout,K,stride,pad", [
    (300, 1, 1, 2048, 256, 1, 1, 0),       # FC-shaped (flat), bn 256, Cout tiles in cluster pairs (TMA-slab epilogue), ragged pixel count
    (1024, 1, 1, 4096, 1024, 1, 1, 0),     # 64 cluster tiles
    (64, 1, 1, 76800, 256, 1, 1, 0),       # 300 cluster tiles on 74 clusters: the slab ring wraps across tiles and accumulators
    (130, 1, 1, 2336, 320, 1, 1, 0),       # TMA-slab epilogue with a partial last Cin tile (2336 = 9*256 + 32) and a partial Cout tile
    (2, 9, 11, 256, 256, 3, 1, 1),         # 3x3, 9 taps, bn 256, cluster pairs: slab columns = tap*Cin + ci
    (64, 1, 1, 96, 136, 1, 1, 0),          # Cin not a multiple of the tile, Cout not a multiple of 128
    (2, 12, 20, 64, 64, 3, 2, 1),          # 3x3 stride 2 (parity maps, 9 taps), bn 64
    (2, 16, 16, 160, 128, 3, 1, 1),        # 3x3 stride 1, Cin = 160 (partial 256-wide tile)
])
def test_wgrad_sgd_epilogue_is_bit_identical_to_wgrad_then_sgd(N, H, W, Cin, Cout, K, stride, pad):
    g = torch.Generator(device=DEV).manual_seed(N * 7 + Cin)
    OH, OW = (H + 2 * pad - K) // stride + 1, (W + 2 * pad - K) // stride + 1
    n = Cout * K * K * Cin
    w_a = torch.randn(n, device=DEV, generator=g)
    w_b = w_a.clone()
    buf_a, buf_b = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    sh_a = torch.zeros(n, device=DEV, dtype=torch.bfloat16)
    sh_b = torch.zeros(n, device=DEV, dtype=torch.bfloat16)
    dw = torch.empty(n, device=DEV)
    desc = F_._conv_desc(N, H, W, Cin, Cout, K, K, stride, pad, "umma_bf16", torch.bfloat16, torch.bfloat16)
    ws = F_.workspace(lib.da_conv_workspace_bytes(ctypes.byref(desc)), torch.device(DEV), "conv")
    for step in range(3):
        x = torch.randn(N, H, W, Cin, device=DEV, generator=g).to(torch.bfloat16)
        dz = torch.randn(N, OH, OW, Cout, device=DEV, generator=g).to(torch.bfloat16)
        check(lib.da_conv_backward_weight(ctypes.byref(desc), F_._ptr(x), F_._ptr(dz), F_._ptr(dw), F_._ptr(ws), ws.numel(), None), "wgrad")
        check(lib.da_sgd_step(F_._ptr(w_a), F_._ptr(dw), F_._ptr(buf_a), n, LR, MU, WD, int(step == 0), F_._ptr(sh_a), None), "sgd")
        rec = _lib.SgdFuse(w_b.data_ptr(), buf_b.data_ptr(), sh_b.data_ptr(), LR, MU, WD, int(step == 0))
        check(lib.da_conv_backward_weight_sgd(ctypes.byref(desc), F_._ptr(x), F_._ptr(dz), ctypes.byref(rec), F_._ptr(ws), ws.numel(), None),
              "wgrad_sgd")
        torch.cuda.synchronize()
        assert torch.equal(w_a, w_b), f"step {step}: master"
        assert torch.equal(buf_a, buf_b), f"step {step}: momentum"
        assert torch.equal(sh_a.view(torch.int16), sh_b.view(torch.int16)), f"step {step}: operand copy"


def test_wgrad_sgd_rejects_unsupported():
    desc = F_._conv_desc(8, 1, 1, 40, 64, 1, 1, 1, 0, "umma_bf16", torch.bfloat16, torch.bfloat16)
    ws = F_.workspace(lib.da_conv_workspace_bytes(ctypes.byref(desc)), torch.device(DEV), "conv")
    t = torch.zeros(64 * 40, device=DEV)
    x = torch.zeros(8, 40, device=DEV, dtype=torch.bfloat16)
    dz = torch.zeros(8, 64, device=DEV, dtype=torch.bfloat16)
    rec = _lib.SgdFuse(t.data_ptr(), t.data_ptr(), None, LR, MU, WD, 0)
    assert lib.da_conv_backward_weight_sgd(ctypes.byref(desc), F_._ptr(x), F_._ptr(dz), ctypes.byref(rec), F_._ptr(ws), ws.numel(), None) != 0
    assert "multiple of 32" in _lib.last_error()
    desc2 = F_._conv_desc(8, 1, 1, 64, 64, 1, 1, 1, 0, "simt_f32", torch.float32, torch.float32)
    assert lib.da_conv_backward_weight_sgd(ctypes.byref(desc2), F_._ptr(x), F_._ptr(dz), ctypes.byref(rec), F_._ptr(ws), ws.numel(), None) != 0


def test_fused_sgd_option_matches_plain_fused_sgd_through_the_layer():
    """optim.FusedSGD(fuse_wgrad=[w]) + functional.dense_layer against the unfused optimizer on a twin layer."""
    uda.set_engine("umma_bf16")
    torch.manual_seed(3)
    lin_a = torch.nn.Linear(1024, 256).to(DEV)
    lin_b = torch.nn.Linear(1024, 256).to(DEV)
    lin_b.load_state_dict(lin_a.state_dict())
    opt_a = optim.FusedSGD(list(lin_a.parameters()), lr=LR, momentum=MU, weight_decay=WD, fuse_wgrad=[lin_a.weight])
    opt_b = optim.FusedSGD(list(lin_b.parameters()), lr=LR, momentum=MU, weight_decay=WD)
    assert all(p is not lin_a.weight for p in opt_a.params)
    g = torch.Generator(device=DEV).manual_seed(11)
    try:
        for step in range(3):
            x = torch.randn(200, 1, 1, 1024, device=DEV, generator=g).to(torch.bfloat16).requires_grad_(True)
            t = torch.randn(200, 1, 1, 256, device=DEV, generator=g)
            grads = []
            for lin in (lin_a, lin_b):
                xi = x.detach().clone().requires_grad_(True)
                y = F_.dense_layer(xi, lin.weight, None, lin.bias, relu=True)
                (y.float() * t).sum().backward()
                grads.append(xi.grad)
            assert lin_a.weight.grad is None                      # never materialised
            assert torch.equal(grads[0], grads[1])                # the data gradient used the pre-update operand copy
            opt_a.step(); opt_a.zero_grad()
            opt_b.step(); opt_b.zero_grad()
            torch.cuda.synchronize()
            assert torch.equal(lin_a.weight.data, lin_b.weight.data), f"step {step}"
            assert torch.equal(lin_a.bias.data, lin_b.bias.data)
            assert torch.equal(F_.bf16_shadow(lin_a.weight).view(torch.int16), F_.bf16_shadow(lin_b.weight).view(torch.int16))
    finally:
        F_.MANAGED_WGRAD.clear()


def test_fused_sgd_state_dict_has_the_torch_sgd_layout_and_moves_both_ways():
    """FusedSGD.state_dict() / load_state_dict() use torch.optim.SGD's layout ({'state': {i: {'momentum_buffer'}}, 'param_groups'}),
    the 'optimizer' section mmcv's save_checkpoint writes for the reference (mmdet/apis/train.py:127): after the same two
    steps both optimizers hold the same momentum; a FusedSGD restored from either state_dict continues exactly like the
    uninterrupted one; a parameter that never received a gradient has no entry (and none is invented on load)."""
    torch.manual_seed(0)
    shapes = [(64, 32), (48,), (16, 8, 3, 3), (5,)]
    base = [torch.randn(*s, device=DEV) for s in shapes]
    grads = [[torch.randn(*s, device=DEV) for s in shapes] for _ in range(4)]

    def make(cls, **kw):
        ps = [torch.nn.Parameter(b.clone()) for b in base]
        return ps, cls(ps, lr=LR, momentum=MU, weight_decay=WD, **kw)

    def run(ps, opt, steps):
        for k in steps:
            for i, p in enumerate(ps):
                p.grad = None if i == 3 else grads[k][i].clone()      # parameter 3 never gets a gradient
            opt.step()

    ps_f, fused = make(optim.FusedSGD, shadow_bf16=False)
    ps_t, ref = make(torch.optim.SGD)
    run(ps_f, fused, (0, 1))
    run(ps_t, ref, (0, 1))
    sd_f, sd_t = fused.state_dict(), ref.state_dict()
    assert set(sd_f) == set(sd_t) == {"state", "param_groups"}
    assert sorted(sd_f["state"]) == sorted(sd_t["state"]) == [0, 1, 2]          # no entry for the parameter without gradients
    for i in sd_t["state"]:
        assert torch.allclose(sd_f["state"][i]["momentum_buffer"], sd_t["state"][i]["momentum_buffer"], rtol=1e-6, atol=1e-7)
    g_f, g_t = sd_f["param_groups"][0], sd_t["param_groups"][0]
    assert g_f["params"] == g_t["params"] and all(g_f[k] == g_t[k] for k in ("lr", "momentum", "weight_decay", "dampening", "nesterov"))
    run(ps_f, fused, (2, 3))                                                     # the uninterrupted run
    for source in (sd_f, sd_t):                                                  # ... and two resumed ones
        ps_r, resumed = make(optim.FusedSGD, shadow_bf16=False)
        with torch.no_grad():
            for p, q in zip(ps_r, ps_t if source is sd_t else ps_t):             # weights after two steps (the torch run's)
                p.copy_(q)
        resumed.load_state_dict(source)
        assert id(ps_r[3]) not in resumed.state
        run(ps_r, resumed, (2, 3))
        for p, q in zip(ps_r, ps_f):
            assert torch.allclose(p, q, rtol=2e-6, atol=1e-6)
