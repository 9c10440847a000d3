"""GPU parity suite (-m gpu): the CUDA path, called through the C-ABI (ctypes -> libda_b200.so),
against the oracle and the committed golden vectors.

Tolerances (BASELINE.json north_star): sampling grids / batch indices bit-exact; fp32 values,
losses and gradients <= 1e-5 relative (to the max magnitude of the reference tensor) on the fp32
engines; the tcgen05 bf16 engine is stated separately (bf16 inputs, fp32 accumulate: 2e-2).
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import unsupervised_domain_adaptation_object_detection_implementation_b200 as uda  # noqa: E402
from unsupervised_domain_adaptation_object_detection_implementation_b200 import functional as F_, ops, da_losses  # noqa: E402
from unsupervised_domain_adaptation_object_detection_implementation_b200.roi_extractors import SingleRoIExtractor, bbox2roi  # noqa: E402
from unsupervised_domain_adaptation_object_detection_implementation_b200 import hotpath  # noqa: E402
from oracle import da_oracle, roi_align as oracle_roi, seeded  # noqa: E402
from helpers import HEADS, build_head, rel_err, check_summary  # noqa: E402

DEV = "cuda"
FP32_TOL = 1e-5          # north_star: fp32 values / losses / gradients; met by simt_f32 AND the tcgen05 engine umma_bf16x6
BF16X3_TOL = 2e-4        # 2-way split (legacy, cheaper): not an fp32-class engine
BF16_TOL = 3e-2


@pytest.fixture(autouse=True)
def _engine_reset():
    yield
    uda.set_engine("umma_bf16")


# ---------------------------------------------------------------------------- RoIAlign
def _roi_case(golden):
    g = golden("roi_align_torchvision.pt")
    return g, g["feat"].to(DEV), g["rois"].to(DEV), 1.0 / g["stride"]


@pytest.mark.parametrize("channels_last", [False, True])
def test_roi_align_forward_backward_vs_golden(golden, channels_last):
    g, feat, rois, scale = _roi_case(golden)
    if channels_last:
        feat = feat.contiguous(memory_format=torch.channels_last)
    feat.requires_grad_(True)
    out, grid = F_.roi_align(feat, rois, 7, scale, 0, True, return_grid=True)
    assert torch.equal(grid.cpu(), g["grid"])                                  # bit-exact sampling grid
    assert rel_err(out, g["out"]) <= FP32_TOL
    (out * g["cot"].to(DEV)).sum().backward()
    assert rel_err(feat.grad, g["dfeat"]) <= FP32_TOL


def test_roi_align_layouts_dtypes_and_module_surface(golden):
    g, feat, rois, scale = _roi_case(golden)
    layer = ops.RoIAlign(output_size=7, spatial_scale=scale, sampling_ratio=0)      # mmcv.ops.RoIAlign signature
    assert layer.output_size == (7, 7) and layer.aligned is True
    out = layer(feat, rois)
    assert rel_err(out, g["out"]) <= FP32_TOL
    out_hwc = F_.roi_align(feat, rois, 7, scale, 0, True, out_layout="rhwc")
    assert rel_err(out_hwc.permute(0, 3, 1, 2), g["out"]) <= FP32_TOL
    # legacy (aligned=False, sampling_ratio=2)
    out2 = F_.roi_align(feat, rois[:24], 7, scale, 2, False)
    assert rel_err(out2, g["out_legacy_sr2"]) <= FP32_TOL
    # bf16 features: compare with the oracle evaluated on the bf16-rounded map
    fb = feat.to(torch.bfloat16)
    ref, _, _ = oracle_roi.roi_align_forward(fb.float().cpu().numpy(), rois.cpu().numpy(), 7, scale)
    # (a) tensor-core path (default for bf16): interpolation weights rounded to bf16, fp32 accumulate
    ob = F_.roi_align(fb, rois, 7, scale, 0, True, out_dtype=torch.float32)
    assert rel_err(ob, torch.from_numpy(ref)) <= 1e-2
    ob16 = F_.roi_align(fb, rois, 7, scale, 0, True)
    assert ob16.dtype == torch.bfloat16 and rel_err(ob16.float(), torch.from_numpy(ref)) <= 1e-2
    # (b) CUDA-core path on the same bf16 features: fp32 weights and accumulation
    F_.set_option("roi_no_tc", 1)
    try:
        oc = F_.roi_align(fb, rois, 7, scale, 0, True, out_dtype=torch.float32)
    finally:
        F_.set_option("roi_no_tc", 0)
    assert rel_err(oc, torch.from_numpy(ref)) <= FP32_TOL


@pytest.mark.parametrize("C,H,W,R,N", [(320, 33, 47, 300, 3), (256, 64, 128, 512, 2), (64, 20, 30, 63, 1), (2048, 64, 128, 1024, 2)])
@pytest.mark.parametrize("out_dtype", [torch.bfloat16, torch.float32])
def test_roi_align_tensor_core_bin_major_layout_matches_reference_layout(C, H, W, R, N, out_dtype):
    """DA_ROI_OUT_RHWC on the tensor-core kernels ([R,7,7,C] RoI tensor).  Forward: the [49][128] tile leaves through one tensor
    store; products and their order are those of the [R,C,7,7] mode, so the values must be BIT-identical to it (which the other
    tests pin to the oracle).  Backward: the gradient operand is fetched MN-major by TMA, and only the bin rows that meet the
    pixel tile are fetched and multiplied; the skipped products are exact zeros but the k-steps group the bins differently, so
    the bf16 result may differ from the [R,C,7,7] mode by the rounding of the last bit: <= 2 bf16 ulp at the largest magnitude
    (both modes are also checked against the oracle below / elsewhere).  Ragged channel counts (C % 256 != 0), adversarial /
    empty RoIs and the bench size."""
    if C == 2048 and out_dtype is torch.float32:
        pytest.skip("bench size is covered in bf16")
    stride = 16
    feat = torch.relu(torch.randn(N, H, W, C, generator=torch.Generator().manual_seed(5))).to(DEV).bfloat16().permute(0, 3, 1, 2)
    rois = torch.cat([seeded.synthetic_rois(R // N, N, H * stride, W * stride, 3),
                      seeded.adversarial_rois(N, H * stride, W * stride)]).to(DEV)
    rois = torch.cat([rois, torch.tensor([[-1.0, 0, 0, 64, 64], [float(N), 8, 8, 99, 99]], device=DEV)])   # not on any image: zeros
    cot = torch.randn(rois.shape[0], C, 7, 7, generator=torch.Generator().manual_seed(6)).to(DEV).bfloat16()
    res = {}
    for layout in ("rchw", "rhwc"):
        x = feat.detach().clone(memory_format=torch.preserve_format).requires_grad_(True)
        out = F_.roi_align(x, rois, 7, 1.0 / stride, out_dtype=out_dtype, out_layout=layout)
        g_in = cot if layout == "rchw" else cot.permute(0, 2, 3, 1).contiguous()
        if layout == "rhwc":
            assert out.shape == (rois.shape[0], 7, 7, C) and out.is_contiguous()
        # bf16 cotangent in the layout of the output -> the tensor-core backward in both cases (fp32 outputs: forward only)
        g = torch.autograd.grad(out, x, g_in)[0] if out_dtype is torch.bfloat16 else None
        res[layout] = (out if layout == "rchw" else out.permute(0, 3, 1, 2), g)
    assert res["rhwc"][0].dtype == out_dtype and float(res["rchw"][0].float().abs().max()) > 0
    assert torch.equal(res["rchw"][0], res["rhwc"][0])
    assert float(res["rchw"][0][-2:].float().abs().max()) == 0.0           # RoIs of no image: zeros in both layouts
    if out_dtype is torch.bfloat16:
        assert res["rchw"][1].dtype == torch.bfloat16 and float(res["rchw"][1].float().abs().max()) > 0
        assert rel_err(res["rhwc"][1].float(), res["rchw"][1].float()) <= 2 * 2.0 ** -8
        if C <= 320:      # and against the oracle's transposed map directly (fp64 loop; small sizes)
            gref = oracle_roi.roi_align_backward(cot.float().cpu().numpy(), rois.cpu().numpy(), (N, C, H, W), 7, 1.0 / stride)
            assert rel_err(res["rhwc"][1].float(), torch.from_numpy(gref)) <= 1e-2


def test_bin_major_hot_path_matches_reference_order_hot_path():
    """hotpath.DAFOrgHotPath(roi_layout="rhwc") against the default layout on the same reference-order state_dict: same losses
    and gradients up to the summation order of the first shared FC; load_state_dict / state_dict speak the reference's
    (channel, bin) column order in both (convfc_bbox_head.py:229), the held weight is the (bin, channel) permutation."""
    uda.set_engine("umma_bf16")
    C, Hf, Wf = 256, 24, 40
    ref_model = hotpath.DAFOrgHotPath(C, 16, 1024).eval()
    seeded.fill_state_(ref_model, 11, "binmajor.")
    sd = ref_model.state_dict()
    alt = hotpath.DAFOrgHotPath(C, 16, 1024, roi_layout="rhwc").eval()
    alt.load_state_dict(sd)
    w_ref, w_alt = sd["bbox_head.shared_fcs.0.weight"], alt.bbox_head.shared_fcs[0].weight.detach()
    assert torch.equal(w_alt.view(1024, 49, C), w_ref.view(1024, C, 49).transpose(1, 2))
    assert all(torch.equal(v, alt.state_dict()[k]) for k, v in sd.items())
    c5 = torch.relu(torch.randn(2, Hf, Wf, C, generator=torch.Generator().manual_seed(1))).to(DEV).bfloat16().permute(0, 3, 1, 2)
    boxes = [seeded.synthetic_rois(96, 1, Hf * 16, Wf * 16, s)[:, 1:].to(DEV) for s in (1, 2)]
    got = {}
    for name, m in (("rchw", ref_model), ("rhwc", alt)):
        m = m.to(DEV)
        x = c5.detach().clone(memory_format=torch.preserve_format).requires_grad_(True)
        losses = m.forward_train(x, boxes, [0, 1])
        total, _ = hotpath.parse_losses(losses)
        total.backward()
        got[name] = ({k: float(v) for k, v in losses.items()}, x.grad.float(),
                     m.bbox_head.to_reference_order(m.bbox_head.shared_fcs[0].weight.grad).float())
    for k, v in got["rchw"][0].items():
        assert abs(got["rhwc"][0][k] - v) <= 2e-3 * abs(v) + 1e-6, (k, v, got["rhwc"][0][k])
    assert rel_err(got["rhwc"][1], got["rchw"][1]) <= BF16_TOL
    assert float((got["rhwc"][2] - got["rchw"][2]).norm() / got["rchw"][2].norm()) <= BF16_TOL


def test_roi_align_edge_cases():
    feat = seeded.seeded_tensor("edge.feat", (2, 6, 9, 11), 0).to(DEV)
    empty = F_.roi_align(feat, torch.zeros(0, 5, device=DEV), 7, 0.25)
    assert empty.shape == (0, 6, 7, 7)
    feat.requires_grad_(True)
    out = F_.roi_align(feat, torch.zeros(0, 5, device=DEV), 7, 0.25)
    out.sum().backward()
    assert float(feat.grad.abs().max()) == 0.0
    # batch index out of range: bounds-checked -> zeros, and raises when validated
    bad = torch.tensor([[2.0, 0, 0, 16, 16], [-1.0, 0, 0, 16, 16], [1.0, 4, 4, 30, 30]], device=DEV)
    out = F_.roi_align(feat.detach(), bad, 7, 0.25)
    assert float(out[:2].abs().max()) == 0.0 and float(out[2].abs().max()) > 0
    with pytest.raises(RuntimeError, match="batch index"):
        F_.roi_align(feat.detach(), bad, 7, 0.25, validate=True)
    with pytest.raises(RuntimeError):
        F_.roi_align(feat.detach(), bad, 5, 0.25)           # only output_size=7 is built


@pytest.mark.parametrize("C,H,W,R", [(320, 33, 47, 300), (256, 64, 128, 512), (2, 5, 7, 40)])
def test_roi_align_vs_oracle_ragged_shapes(C, H, W, R):
    N, stride = 3, 8
    feat = seeded.seeded_tensor("rag.feat", (N, C, H, W), 2)
    rois = torch.cat([seeded.synthetic_rois(R // N, N, H * stride, W * stride, 2, 6.0, 400.0),
                      seeded.adversarial_rois(N, H * stride, W * stride)])
    ref, grid_ref, _ = oracle_roi.roi_align_forward(feat.numpy(), rois.numpy(), 7, 1.0 / stride, threads=8)
    f = feat.to(DEV).requires_grad_(True)
    out, grid = F_.roi_align(f, rois.to(DEV), 7, 1.0 / stride, 0, True, return_grid=True)
    assert np.array_equal(grid.cpu().numpy(), grid_ref)
    assert rel_err(out, torch.from_numpy(ref)) <= FP32_TOL
    cot = seeded.seeded_tensor("rag.cot", tuple(out.shape), 2)
    (out * cot.to(DEV)).sum().backward()
    gref = oracle_roi.roi_align_backward(cot.numpy(), rois.numpy(), (N, C, H, W), 7, 1.0 / stride)
    assert rel_err(f.grad, torch.from_numpy(gref)) <= FP32_TOL


def test_roi_align_full_size_properties():
    """BASELINE config 4 size (4 x 2048 x 64 x 128, 2048 RoIs): size-independent properties."""
    N, C, H, W, R = 4, 2048, 64, 128, 2048
    rois = seeded.synthetic_rois(R // N, N, H * 16, W * 16, 0).to(DEV)
    g = torch.Generator(device=DEV).manual_seed(0)
    f1 = torch.randn(N, H, W, C, device=DEV, generator=g).permute(0, 3, 1, 2)   # channels_last storage
    f2 = torch.randn(N, H, W, C, device=DEV, generator=g).permute(0, 3, 1, 2)
    o1, o2 = F_.roi_align(f1, rois, 7, 1 / 16), F_.roi_align(f2, rois, 7, 1 / 16)
    o12 = F_.roi_align(2.0 * f1 - f2, rois, 7, 1 / 16)
    assert rel_err(o12, 2.0 * o1 - o2) <= 1e-5                                 # linearity
    ones = F_.roi_align(torch.ones_like(f1), rois, 7, 1 / 16)                   # RoIs lie inside the image
    assert float((ones - 1).abs().max()) <= 1e-5                               # partition of unity
    # adjointness <A f, g> == <f, A^T g>: ties backward to forward at full size
    f1 = f1.detach().requires_grad_(True)
    cot = torch.randn(R, C, 7, 7, device=DEV, generator=g)
    out = F_.roi_align(f1, rois, 7, 1 / 16)
    lhs = (out.double() * cot.double()).sum()
    out.backward(cot)
    rhs = (f1.grad.double() * f1.detach().double()).sum()
    assert abs(float(lhs - rhs)) <= 1e-6 * abs(float(lhs))
    # a slice against the oracle
    sub = rois[:16].clone()
    sub[:, 0] = 0
    ref, _, _ = oracle_roi.roi_align_forward(f1[:1, :64].detach().cpu().numpy(), sub.cpu().numpy(), 7, 1 / 16)
    got = F_.roi_align(f1[:1, :64].detach().contiguous(), sub, 7, 1 / 16)
    assert rel_err(got, torch.from_numpy(ref)) <= FP32_TOL


def test_roi_align_tensor_core_path_full_size():
    """bf16 tensor-core forward at BASELINE config 4 size (+ adversarial RoIs that fall back to the CUDA-core
    kernel): against the CUDA-core bf16 path everywhere, and against the oracle on a slice."""
    N, C, H, W, R = 4, 2048, 64, 128, 2048
    rois = torch.cat([seeded.synthetic_rois(R // N, N, H * 16, W * 16, 0), seeded.adversarial_rois(N, H * 16, W * 16)]).to(DEV)
    g = torch.Generator(device=DEV).manual_seed(1)
    f = torch.relu(torch.randn(N, H, W, C, device=DEV, generator=g)).to(torch.bfloat16).permute(0, 3, 1, 2)
    tc = F_.roi_align(f, rois, 7, 1 / 16)
    F_.set_option("roi_no_tc", 1)
    try:
        cc = F_.roi_align(f, rois, 7, 1 / 16)
    finally:
        F_.set_option("roi_no_tc", 0)
    assert tc.dtype == torch.bfloat16 and tc.shape == (rois.shape[0], C, 7, 7)
    assert rel_err(tc.float(), cc.float()) <= 1e-2
    sub = rois[-40:].clone()
    ref, _, _ = oracle_roi.roi_align_forward(f[:, :96].float().cpu().numpy(), sub.cpu().numpy(), 7, 1 / 16)
    got = F_.roi_align(f[:, :96].contiguous(memory_format=torch.channels_last), sub, 7, 1 / 16, out_dtype=torch.float32)
    assert rel_err(got, torch.from_numpy(ref)) <= 1e-2


@pytest.mark.parametrize("N,C,H,W,R", [(3, 328, 37, 53, 1500), (2, 8, 20, 30, 64), (1, 512, 16, 16, 1100)])
def test_roi_align_tensor_core_backward_vs_oracle(N, C, H, W, R):
    """bf16 gradients take the tcgen05 gather kernel (TMEM-resident tile accumulators).  Ragged shapes: channel counts
    that are not a multiple of the 256-channel CTA slice, partial 16x16 pixel tiles, more RoIs than one list chunk
    (1024), adversarial boxes.  Reference: the fp64 oracle backward on the SAME bf16-rounded cotangent."""
    stride = 8
    rois = torch.cat([seeded.synthetic_rois(R // N, N, H * stride, W * stride, 5, 6.0, 300.0),
                      seeded.adversarial_rois(N, H * stride, W * stride)])
    feat = seeded.seeded_tensor("tcb.feat", (N, C, H, W), 5).to(torch.bfloat16)
    f = feat.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    out = F_.roi_align(f, rois.to(DEV), 7, 1.0 / stride)
    assert out.dtype == torch.bfloat16
    cot = seeded.seeded_tensor("tcb.cot", tuple(out.shape), 5).to(torch.bfloat16)
    (g_tc,) = torch.autograd.grad(out, f, cot.to(DEV), retain_graph=True)
    assert g_tc.dtype == torch.bfloat16 and g_tc.shape == f.shape
    gref = torch.from_numpy(oracle_roi.roi_align_backward(cot.float().numpy(), rois.numpy(), (N, C, H, W), 7, 1.0 / stride))
    # bf16 tap weights (2^-9) and one bf16 rounding of the result
    assert rel_err(g_tc.float(), gref) <= 1e-2
    F_.set_option("roi_no_tc", 1)         # CUDA-core kernel on the same inputs: fp32 weights, fp32 result
    try:
        (g_cc,) = torch.autograd.grad(out, f, cot.to(DEV))
    finally:
        F_.set_option("roi_no_tc", 0)
    assert rel_err(g_cc.float(), gref) <= 4e-3
    # pixels no RoI touches are exactly zero in both
    untouched = gref == 0
    if bool(untouched.any()):
        assert float(g_tc.float()[untouched.to(DEV)].abs().max()) == 0.0


def test_single_roi_extractor_and_level_mapping():
    ext = SingleRoIExtractor(dict(type="RoIAlign", output_size=7, sampling_ratio=0), out_channels=16, featmap_strides=[4, 8, 16, 32])
    feats = [seeded.seeded_tensor(f"fpn{i}", (2, 16, 128 >> i, 192 >> i), 0).to(DEV) for i in range(4)]
    boxes = [seeded.synthetic_rois(50, 1, 512, 768, s, 8.0, 700.0)[:, 1:] for s in (0, 1)]
    rois = bbox2roi([b.to(DEV) for b in boxes])
    lv = ext.map_roi_levels(rois, 4)
    assert np.array_equal(lv.cpu().numpy(), oracle_roi.map_roi_levels(rois.cpu().numpy(), 4, 56.0))   # bit-exact
    out = ext(feats, rois)
    ref = np.zeros((rois.shape[0], 16, 7, 7), np.float32)
    for i in range(4):
        idx = (lv == i).nonzero().squeeze(1).cpu().numpy()
        if len(idx):
            ref[idx] = oracle_roi.roi_align_forward(feats[i].cpu().numpy(), rois.cpu().numpy()[idx], 7, 1.0 / (4 << i))[0]
    assert rel_err(out, torch.from_numpy(ref)) <= FP32_TOL
    single = SingleRoIExtractor(dict(type="RoIAlign", output_size=7, sampling_ratio=0), out_channels=16, featmap_strides=[16])
    assert single([feats[2]], rois[:0]).shape == (0, 16, 7, 7)


# ---------------------------------------------------------------------------- losses
def _grad_pair(fn_cuda, fn_ref, tensors):
    cu = [t.to(DEV).float().requires_grad_(True) for t in tensors]
    rf = [t.double().requires_grad_(True) for t in tensors]
    a, b = fn_cuda(*cu), fn_ref(*rf)
    a.backward(torch.tensor(0.7, device=DEV))
    (b * 0.7).backward()
    return a, b, cu, rf


@pytest.mark.parametrize("gt", [[0, 1], [1, 0], [0, 0], [1, 1], [0, 1, 1, 0]])
@pytest.mark.parametrize("whole", [True, False])
def test_pixel_domain_loss(gt, whole):
    n = len(gt)
    p = seeded.seeded_tensor("pl.p", (n, 1, 37, 53), 0, scale=2.0)
    gtt = torch.tensor(gt)
    ref_fn = da_oracle.daf_image_loss if whole else da_oracle.patch_loss
    a, b, cu, rf = _grad_pair(lambda x: F_.pixel_domain_loss(x, gtt.to(DEV), whole), lambda x: ref_fn(x, gtt), [p])
    assert abs(float(a) - float(b)) <= FP32_TOL * abs(float(b))
    assert rel_err(cu[0].grad, rf[0].grad) <= FP32_TOL


def test_pixel_domain_loss_large_map():
    p = seeded.seeded_tensor("pl.big", (2, 1, 128, 256), 1, scale=3.0)
    gtt = torch.tensor([0, 1])
    for whole, ref_fn in ((True, da_oracle.daf_image_loss), (False, da_oracle.patch_loss)):
        a = F_.pixel_domain_loss(p.to(DEV), gtt.to(DEV), whole)
        assert abs(float(a) - float(ref_fn(p.double(), gtt))) <= FP32_TOL * float(a)


@pytest.mark.parametrize("on_sigmoid", [False, True])
@pytest.mark.parametrize("R", [2, 1024, 1500])
def test_ce2(on_sigmoid, R):
    z = seeded.seeded_tensor("ce.z", (R, 2), 0, scale=2.0)
    lab = (seeded.seeded_tensor("ce.l", (R,), 0) > 0).long()
    ref = lambda x: da_oracle.ce2(torch.sigmoid(x) if on_sigmoid else x, lab)
    a, b, cu, rf = _grad_pair(lambda x: F_.ce2(x, lab.to(DEV), on_sigmoid)[0], ref, [z])
    assert abs(float(a) - float(b)) <= FP32_TOL * abs(float(b))
    assert rel_err(cu[0].grad, rf[0].grad) <= FP32_TOL
    if on_sigmoid:
        _, pred = F_.ce2(z.to(DEV), lab.to(DEV), True)
        assert rel_err(pred, torch.sigmoid(z)) <= 1e-6


def test_ce2_known_answer():
    loss, _ = F_.ce2(torch.tensor([[100.0, -100.0]], device=DEV), torch.tensor([1], device=DEV), False)
    assert abs(float(loss) - 200.0) < 1e-3     # reference tests/test_metrics/test_losses.py:18-33


def test_focal_vs_reference_golden(golden):
    g = golden("focal_loss.pt")
    u = g["u"].to(DEV).requires_grad_(True)
    loss = da_losses.FocalLoss(gamma=2.0, alpha=0.25)(u, g["labels"].to(DEV))
    assert abs(float(loss) - float(g["loss"])) <= FP32_TOL * float(g["loss"])
    loss.backward()
    assert rel_err(u.grad, g["du"]) <= FP32_TOL
    with pytest.raises(NotImplementedError):
        da_losses.FocalLoss(use_sigmoid=False)


@pytest.mark.parametrize("R", [0, 30, 1024])
def test_consistency_loss(R):
    img = seeded.feature_map("cs.img", (2, 1, 32, 64), 0)
    pred = torch.sigmoid(seeded.seeded_tensor("cs.pred", (R, 2), 0))
    lab = torch.cat([torch.zeros(R // 2), torch.ones(R - R // 2)]).long()
    a, b, cu, rf = _grad_pair(lambda i, p: F_.consistency_loss(i, p, lab.to(DEV)),
                              lambda i, p: da_oracle.consistency_loss(i, p, lab), [img, pred])
    assert abs(float(a) - float(b)) <= FP32_TOL * max(abs(float(b)), 1e-6)
    if R:
        assert rel_err(cu[0].grad, rf[0].grad) <= 1e-4      # sign() sum of +-1 terms, fp32 mean
        assert rel_err(cu[1].grad, rf[1].grad) <= FP32_TOL
        if R <= 64:
            c = da_oracle.consistency_loss_loop(img.double(), pred.double(), lab)
            assert abs(float(a) - float(c)) <= FP32_TOL * abs(float(c))


def test_grl_and_softmax_dim0():
    x = seeded.seeded_tensor("grl.x", (5, 7), 0).to(DEV).requires_grad_(True)
    (F_.gradient_scalar(x, -0.3) * 2.0).sum().backward()
    assert torch.allclose(x.grad, torch.full_like(x, -0.6))
    s = seeded.seeded_tensor("sm.s", (96, 96), 0, scale=3.0)
    a = s.to(DEV).requires_grad_(True)
    b = s.double().requires_grad_(True)
    pa, pb = F_.softmax_dim0(a), torch.softmax(b, dim=0)
    assert rel_err(pa, pb) <= FP32_TOL
    cot = seeded.seeded_tensor("sm.c", (96, 96), 0)
    (pa * cot.to(DEV)).sum().backward()
    (pb * cot.double()).sum().backward()
    assert rel_err(a.grad, b.grad) <= FP32_TOL


# ---------------------------------------------------------------------------- dense engines
CONV_CASES = [
    # N, H, W, Cin, Cout, k, stride, pad
    (2, 9, 13, 64, 128, 1, 1, 0),      # flat 1x1
    (2, 9, 13, 64, 72, 1, 1, 1),       # 1x1 with padding (SRM conv1, Q12)
    (2, 10, 14, 64, 136, 3, 1, 1),
    (2, 11, 15, 128, 64, 3, 2, 1),     # stride 2, odd extent (Global heads)
    (1, 8, 12, 64, 192, 3, 1, 3),      # padding 3 (SRM conv2)
    (3, 1, 1, 192, 256, 1, 1, 0),      # FC
    (2, 7, 7, 64, 64, 3, 2, 1),        # RoI-conv flavour (local_da.py)
]


def _conv_ref(x, w, stride, pad):
    return torch.nn.functional.conv2d(x.double(), w.double(), None, stride, pad)


@pytest.mark.parametrize("engine,tol", [("simt_f32", FP32_TOL), ("umma_bf16x6", FP32_TOL), ("umma_bf16x3", BF16X3_TOL),
                                        ("umma_bf16", BF16_TOL)])
@pytest.mark.parametrize("case", CONV_CASES)
def test_dense_layer_engines(engine, tol, case):
    N, H, W, Cin, Cout, k, stride, pad = case
    x = seeded.seeded_tensor("cv.x", (N, Cin, H, W), 0)
    w = seeded.seeded_tensor("cv.w", (Cout, Cin, k, k), 0, scale=(Cin * k * k) ** -0.5)
    scale = seeded.seeded_tensor("cv.s", (Cout,), 0, "uniform")
    shift = seeded.seeded_tensor("cv.t", (Cout,), 0, scale=0.1)
    if engine == "umma_bf16":
        # bf16 engine: the reference sees the same bf16-rounded operands (otherwise ReLU masks of
        # pre-activations within bf16 rounding of zero flip and dominate the gradient error)
        x, w = x.bfloat16().float(), w.bfloat16().float()
    xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
    sr, tr = scale.double().requires_grad_(True), shift.double().requires_grad_(True)
    yr = torch.relu(_conv_ref(xr, wr, stride, pad) * sr.view(1, -1, 1, 1) + tr.view(1, -1, 1, 1))
    cot = seeded.seeded_tensor("cv.c", tuple(yr.shape), 0)
    (yr * cot.double()).sum().backward()

    dt = F_.act_dtype(engine)
    xc = x.to(DEV).requires_grad_(True)
    wc = w.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    sc, tc = scale.to(DEV).requires_grad_(True), shift.to(DEV).requires_grad_(True)
    y = F_.dense_layer(F_.to_nhwc(xc, dt), wc, sc, tc, stride=stride, pad=pad, relu=True, engine=engine, grl=-1.0)
    y_nchw = y.permute(0, 3, 1, 2).float()
    assert rel_err(y_nchw, yr) <= tol
    (y_nchw * cot.to(DEV)).sum().backward()
    assert rel_err(xc.grad, -xr.grad) <= tol * 3          # GRL folded into the data gradient
    assert rel_err(wc.grad, wr.grad) <= tol * 3
    assert rel_err(tc.grad, tr.grad) <= tol * 3
    assert rel_err(sc.grad, sr.grad) <= tol * 10


@pytest.mark.parametrize("M,K,Nc", [(1024, 1024, 1024),   # short K, few tiles: 128x64 tiles
                                    (300, 192, 72),        # ragged rows and columns, odd pixel-tile count
                                    (512, 2048, 512),      # CTA-pair MMA (cta_group::2), long K
                                    (256, 4096, 264),      # pair MMA with a ragged last column tile
                                    (256, 16384, 1024),    # 512-wide pair tiles in the forward (>= 256 k-steps), split-K
                                    (1920, 16384, 512),    # ... with an ODD pixel-tile count (phantom tile in the last pair)
                                    (8192, 512, 256),      # 512-wide pair tiles in the weight gradient (>= 128 pixel steps)
                                    (130, 64, 8)])         # single 128x64 tile + 2 rows
def test_fc_tile_variants_bf16(M, K, Nc):
    """Every tile / cluster variant of the tcgen05 GEMM (forward, data gradient, weight gradient) against fp64 on the
    same bf16-rounded operands."""
    uda.set_engine("umma_bf16")
    x = seeded.seeded_tensor("tv.x", (M, K), 3).to(torch.bfloat16)
    w = seeded.seeded_tensor("tv.w", (Nc, K), 3, scale=K ** -0.5).to(torch.bfloat16)
    cot = seeded.seeded_tensor("tv.c", (M, Nc), 3).to(torch.bfloat16)
    xd = x.to(DEV).view(M, 1, 1, K).requires_grad_(True)
    wd = w.float().to(DEV).requires_grad_(True)          # fp32 master, bf16 shadow inside
    y = F_.dense_layer(xd, wd).view(M, Nc)
    y.backward(cot.to(DEV))
    X, Wm, Cm = x.double(), w.double(), cot.double()
    assert rel_err(y.float(), (X @ Wm.t()).float()) <= 1e-2
    assert rel_err(xd.grad.float().view(M, K), (Cm @ Wm).float()) <= 1e-2
    assert rel_err(wd.grad, (Cm.t() @ X).float()) <= 1e-3      # fp32 result


@pytest.mark.parametrize("case", [(2, 64, 64, 512, 256, 3, 1, 1),     # weight gradient: 512-wide pair tiles through the 4-D (tap) loads
                                  (1, 32, 48, 512, 2048, 3, 1, 1),    # data gradient: 512-wide pair tiles, MN-major weights, 288 k-steps
                                  (1, 40, 48, 512, 2048, 3, 1, 1),    # ... with an odd pixel-tile count (15)
                                  (2, 64, 96, 512, 1024, 3, 2, 1),    # stride 2: forward + per-parity-class data gradients on wide tiles
                                  (2, 64, 64, 2048, 512, 1, 1, 0)])   # image-head shape: 32 k-steps in the forward
def test_conv_wide_pair_tiles_bf16(case):
    """The 512-wide CTA-pair tiles (one TMEM accumulator, two N = 256 pair MMAs per k-substep) of the forward / data-gradient
    kernel and of the weight-gradient kernel on 3x3 convolutions, against fp64 on the same bf16-rounded operands."""
    N, H, W, Cin, Cout, k, stride, pad = case
    uda.set_engine("umma_bf16")
    x = seeded.seeded_tensor("wt.x", (N, Cin, H, W), 0).bfloat16().float()
    w = seeded.seeded_tensor("wt.w", (Cout, Cin, k, k), 0, scale=(Cin * k * k) ** -0.5).bfloat16().float()
    xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
    torch.set_num_threads(os.cpu_count() or 8)
    yr = _conv_ref(xr, wr, stride, pad)
    cot = seeded.seeded_tensor("wt.c", tuple(yr.shape), 0).bfloat16().float()
    (yr * cot.double()).sum().backward()
    xc = x.to(DEV).requires_grad_(True)
    wc = w.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    y = F_.dense_layer(F_.to_nhwc(xc, torch.bfloat16), wc, None, None, stride=stride, pad=pad, relu=False, engine="umma_bf16")
    y_nchw = y.permute(0, 3, 1, 2).float()
    assert rel_err(y_nchw, yr) <= 1e-2
    (y_nchw * cot.to(DEV)).sum().backward()
    assert rel_err(xc.grad, xr.grad) <= 1e-2
    assert rel_err(wc.grad, wr.grad) <= 2e-3


@pytest.mark.parametrize("engine", ["simt_f32", "umma_bf16x6", "umma_bf16x3"])
def test_dense_layer_dropout_mask_is_exported(engine):
    N, H, W, Cin, Cout = 2, 6, 8, 64, 128
    x = seeded.seeded_tensor("dr.x", (N, H, W, Cin), 0).to(DEV).requires_grad_(True)
    w = seeded.seeded_tensor("dr.w", (Cout, Cin, 1, 1), 0, scale=Cin ** -0.5).to(DEV).requires_grad_(True)
    seed = 1234567
    y = F_.dense_layer(x, w, relu=True, drop_p=0.5, seed=seed, engine=engine)
    keep = F_.dropout_keep_mask(seed, (N, H, W, Cout), 0.5, DEV)
    assert 0.4 < float(keep.float().mean()) < 0.6
    ref = torch.relu(x.detach().double() @ w.detach().double().view(Cout, Cin).t()) * keep.double() * 2.0
    tol = BF16X3_TOL if engine == "umma_bf16x3" else FP32_TOL
    assert rel_err(y, ref) <= tol
    y.sum().backward()
    xr = x.detach().double().requires_grad_(True)
    (torch.relu(xr @ w.detach().double().view(Cout, Cin).t()) * keep.double() * 2.0).sum().backward()
    assert rel_err(x.grad, xr.grad) <= tol * 3


# ---------------------------------------------------------------------------- heads vs reference golden
@pytest.mark.parametrize("engine,tol", [("simt_f32", FP32_TOL), ("umma_bf16x6", FP32_TOL), ("umma_bf16x3", BF16X3_TOL)])
@pytest.mark.parametrize("name", sorted(HEADS))
def test_heads_match_reference_golden(golden, name, engine, tol):
    g = golden(f"head_{name}.pt")
    uda.set_engine(engine)
    m = build_head(name, g["seed"]).to(DEV)
    x = g["x"].to(DEV).requires_grad_(True)
    out = m(x)
    out = out[0] if isinstance(out, tuple) else out
    assert out.shape == g["out"][0].shape
    assert rel_err(out.float(), g["out"][0]) <= tol
    (out.float() * g["cot"][0].to(DEV)).sum().backward()
    gtol = tol * 5
    if engine == "umma_bf16x3" and name == "roi_local_alignment":
        # three 9216-term reductions in a row: a 1e-5 perturbation of a pre-activation flips a few ReLU masks, which
        # moves the gradient by whole terms (the fp32 engine above matches the reference's own class to 1e-5)
        gtol = 3e-2
    assert rel_err(x.grad, g["dx"]) <= gtol
    params = dict(m.named_parameters())
    for k, summ in g["dparams"].items():
        assert params[k].grad is not None, k
        assert check_summary(params[k].grad, summ, gtol) <= max(gtol, 2e-5), k
    for k in g["no_grad_params"]:            # dead branch (Q10) / unused BN (Q9): no gradient, as in the reference
        assert params[k].grad is None, k


@pytest.mark.parametrize("name", sorted(HEADS))
def test_heads_bf16_engine_vs_bf16_emulating_oracle(golden, name):
    """tcgen05 bf16 engine (the throughput mode), stated separately: the oracle is evaluated with the
    SAME bf16 storage points (operands and stored activations rounded to bf16, fp32/fp64 math in
    between).  Tolerance 3e-2 of the tensor's max magnitude for values, 6e-2 for the input gradient
    (dy and dz are additionally rounded to bf16 in the CUDA backward), 5e-2 relative Frobenius error
    for weight gradients."""
    g = golden(f"head_{name}.pt")
    uda.set_engine("umma_bf16")
    m = build_head(name, g["seed"])
    sd = {k: v.detach().double().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in m.state_dict().items()}
    xr = g["x"].double().requires_grad_(True)
    ref = HEADS[name][1](xr, sd, q="bf16")
    (ref * g["cot"][0].double()).sum().backward()
    m = m.to(DEV)
    x = g["x"].to(DEV).requires_grad_(True)
    out = m(x)
    out = out[0] if isinstance(out, tuple) else out
    assert rel_err(out.float(), ref) <= BF16_TOL
    (out.float() * g["cot"][0].to(DEV)).sum().backward()
    assert rel_err(x.grad, xr.grad) <= 2 * BF16_TOL
    for k, p in m.named_parameters():
        if sd[k].grad is not None and p.grad is not None and k.endswith("weight") and p.dim() >= 2:
            # weight gradients deep in the stack accumulate bf16 rounding of dy/dz at every layer:
            # bounded in the Frobenius norm
            a, b = p.grad.double().cpu(), sd[k].grad.double()
            assert float((a - b).norm() / b.norm()) <= 5e-2, k


def test_head_dropout_training_mode_statistics():
    uda.set_engine("simt_f32")
    m = build_head("local_alignment").to(DEV).train()
    assert not m.bn1.training and not m.bn2.training          # Q9: BN stays in eval
    x = seeded.feature_map("x.local", (2, 64, 6, 10)).to(DEV)
    a, b = m(x), m(x)
    assert float((a - b).abs().max()) > 0                     # dropout active, fresh seed per call
    m.eval()
    assert float((m(x) - m(x)).abs().max()) == 0.0


def test_daf_org_composite_losses_vs_oracle():
    """H1 + L1 + I1 + L4 + L7 with the reference's lambda weights (DAFaster_rcnn_Orig.py:143-157)."""
    uda.set_engine("simt_f32")
    img_head, ins_head = build_head("img_alignment").to(DEV), build_head("instance_alignment").to(DEV)
    c5 = seeded.feature_map("cmp.c5", (2, 64, 8, 12), 0)
    feats = [seeded.feature_map("cmp.f0", (20, 1024), 0), seeded.feature_map("cmp.f1", (28, 1024), 0)]
    gt = torch.tensor([0, 1])
    sd_img = {k: v.cpu().double() for k, v in img_head.state_dict().items()}
    sd_ins = {k: v.cpu().double() for k, v in ins_head.state_dict().items()}
    c5r = c5.double().requires_grad_(True)
    fr = [f.double().requires_grad_(True) for f in feats]
    ref, _, _ = da_oracle.daf_org_da_losses(c5r, fr, gt, sd_img, sd_ins)
    sum(ref.values()).backward()

    c5c = c5.to(DEV).requires_grad_(True)
    fc = [f.to(DEV).requires_grad_(True) for f in feats]
    img_feat = img_head(c5c)
    labels = torch.cat([torch.zeros(20), torch.ones(28)]).long().to(DEV)
    l4, pred = da_losses.instance_ce_loss(ins_head.forward_logits(torch.cat(fc, 0)), labels)
    got = dict(globle_da_loss=0.1 * da_losses.daf_image_loss(img_feat, gt.to(DEV)), local_da_loss=0.1 * l4,
               consistency_loss=0.1 * da_losses.consistency_loss(img_feat, pred, labels))
    for k in ref:
        assert abs(float(got[k]) - float(ref[k])) <= FP32_TOL * abs(float(ref[k])), k
    sum(got.values()).backward()
    assert rel_err(c5c.grad, c5r.grad) <= 5e-5
    assert rel_err(fc[0].grad, fr[0].grad) <= 5e-5 and rel_err(fc[1].grad, fr[1].grad) <= 5e-5


def test_fused_sgd_matches_torch_sgd_and_refreshes_shadow():
    from unsupervised_domain_adaptation_object_detection_implementation_b200 import optim
    torch.manual_seed(0)
    shapes = [(64, 32, 3, 3), (129,), (40, 70), (2 * 65536 + 3,), (3,)]   # multi-chunk tensor with a tail, tiny tensor
    ps = [torch.randn(*s, device=DEV).requires_grad_(True) for s in shapes]
    ps[0].data = ps[0].data.contiguous(memory_format=torch.channels_last)
    qs = [p.detach().clone().requires_grad_(True) for p in ps]
    a = optim.FusedSGD(ps, lr=0.05, momentum=0.9, weight_decay=5e-4)
    b = torch.optim.SGD(qs, lr=0.05, momentum=0.9, weight_decay=5e-4)
    for it in range(3):
        for p, q in zip(ps, qs):
            g = torch.randn_like(q)
            p.grad, q.grad = g.clone(), g.clone()
        a.step()
        b.step()
    for p, q in zip(ps, qs):
        assert rel_err(p, q) <= 1e-6
    sh = F_.bf16_shadow(ps[0])
    assert sh.dtype == torch.bfloat16 and sh.stride() == ps[0].stride()
    assert rel_err(sh.float(), ps[0]) <= 4e-3          # refreshed in the update kernel, no re-cast


# ---------------------------------------------------------------------------- detectors: forward_train surface
def _det_cfg(det_type, backbone_type):
    cfg = uda.Config.fromfile(os.path.join(os.path.dirname(os.path.abspath(__file__)), "fixtures", "cfg", "experiment.py"))
    cfg.model["type"] = det_type
    cfg.model.backbone["type"] = backbone_type
    return cfg


@pytest.mark.parametrize("det,bb,keys", [
    ("DAFasterRCNN_Org", "ResNet_DAF", {"local_da_loss", "globle_da_loss", "consistency_loss"}),
    ("DAFasterRCNN", "ResNet_DA_CBAM", {"local_da_loss", "globle_da_loss", "patch_bottom_loss"}),
    ("MAFasterRCNN", "ResNet_DA", {"local_da_loss", "globle_da_loss"}),
    ("DAFasterRCNN_Deep", "ResNet_DA_Deep", {"local_da_loss", "globle_da_loss", "patch_bottom_loss"}),
])
def test_detectors_forward_train_losses_dict(det, bb, keys):
    """build_detector(cfg.model) -> forward_train(img, img_metas, gt_bboxes, gt_labels, gt_da=...) returns the
    reference's loss keys; train_step gives {loss, log_vars, num_samples}; backward reaches the
    DA heads and the trunk through the GRL."""
    torch.manual_seed(0)
    cfg = _det_cfg(det, bb)
    model = uda.build_detector(cfg.model)
    # the reference's N(0, 0.001) head init leaves ReLU-terminated heads with zero gradients (Q17): use O(1) weights
    seeded.fill_state_(model.backbone.da_head_top, 0, "det.top.")
    model = model.to(DEV).train()
    img = torch.randn(2, 3, 128, 192, device=DEV)
    metas = [dict(img_shape=(128, 192, 3), pad_shape=(128, 192, 3), scale_factor=1.0, flip=False) for _ in range(2)]
    gtb = [torch.tensor([[20., 30., 90., 100.], [100., 20., 180., 110.]], device=DEV), torch.tensor([[40., 40., 120., 120.]], device=DEV)]
    gtl = [torch.tensor([0, 2], device=DEV), torch.tensor([1], device=DEV)]
    losses = model.forward_train(img, metas, gtb, gtl, gt_da=[0, 1])
    assert keys <= set(losses), set(losses)
    assert {"loss_rpn_cls", "loss_rpn_bbox", "loss_cls", "loss_bbox", "acc"} <= set(losses)
    out = model.train_step(dict(img=img, img_metas=metas, gt_bboxes=gtb, gt_labels=gtl, gt_da=[0, 1]), None)
    assert set(out) == {"loss", "log_vars", "num_samples"} and out["num_samples"] == 2
    assert torch.isfinite(out["loss"])
    out["loss"].backward()
    g = model.backbone.da_head_top.conv1.weight.grad
    assert g is not None and torch.isfinite(g).all() and float(g.abs().sum()) > 0
    assert model.backbone.layer4[0].conv1.weight.grad is not None          # reversed gradient reaches the trunk
    unused = {id(p) for p in model.unused_parameters()}
    for n, p in model.named_parameters():
        if id(p) in unused:
            assert p.grad is None, n


# ---------------------------------------------------------------------------- L5 group_local_da_loss
@pytest.mark.parametrize("engine,tol", [("simt_f32", 2e-5), ("umma_bf16x6", 2e-5), ("umma_bf16x3", 3e-4), ("umma_bf16", 3e-2)])
def test_group_local_da_loss_matches_reference_methods(golden, engine, tol):
    """da_losses.group_local_da_loss against the values returned by the reference's own detector methods
    (DAFaster_rcnn.py / MAFaster_rcnn.py / DAFaster_rcnn_Deep.py, run on CPU by oracle/make_golden.group_loss_cases)."""
    from helpers import group_case, group_heads
    uda.set_engine(engine)
    g = golden("group_local_da_loss.pt")
    for name, rec in g.items():
        feats, cls = group_case(rec, name)
        fore, back = group_heads(rec, name)
        fore, back = fore.to(DEV), back.to(DEV)
        torch.manual_seed(rec["rng_seed"])          # centroid draws: the reference's order, from the global CPU generator
        draw = lambda dim, dev: torch.stack([torch.randn([dim]) for _ in range(10)], 0).to(dev)
        val = da_losses.group_local_da_loss([f.to(DEV) for f in feats], [c.to(DEV) for c in cls], fore, back, rec["flavour"],
                                            draw_centroids=draw)
        assert val.dim() == 0 and not val.requires_grad
        assert abs(float(val) - rec["loss"]) <= tol * max(1.0, abs(rec["loss"])), (name, engine, float(val), rec["loss"])


# ---------------------------------------------------------------------------- tools/DA_train.py (entry point surface)
def test_da_train_entry_point_trains_checkpoints_and_resumes(tmp_path):
    """config -> build_detector -> pair sampler -> train_step -> FusedSGD -> checkpoint, then --resume-from."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cfg = os.path.join(root, "tests", "fixtures", "cfg", "experiment.py")
    base = [sys.executable, os.path.join(root, "tools", "DA_train.py"), cfg, "--synthetic", "4", "--img-size", "128x192",
            "--work-dir", str(tmp_path), "--cfg-options", "optimizer.lr=0.0005"]
    out = subprocess.run(base + ["--iters", "3"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-1500:]
    assert "iter 3/3" in out.stdout and "domains [0, 1]" in out.stdout and "globle_da_loss" in out.stdout
    ck = torch.load(tmp_path / "iter_3.pth", map_location="cpu", weights_only=False)
    assert ck["meta"]["iter"] == 3 and "backbone.da_head_top.conv1.weight" in ck["state_dict"]
    # optimizer section in torch.optim.SGD's layout (what mmcv saves for the reference), epoch in the meta
    assert set(ck["optimizer"]) == {"state", "param_groups"} and len(ck["optimizer"]["state"]) > 0
    assert all("momentum_buffer" in v for v in ck["optimizer"]["state"].values()) and "epoch" in ck["meta"]
    out = subprocess.run(base + ["--iters", "5", "--resume-from", str(tmp_path / "iter_3.pth")], capture_output=True, text=True,
                         timeout=600, cwd=root)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-1500:]
    assert "iter 4/5" in out.stdout and "iter 5/5" in out.stdout and "iter 3/5" not in out.stdout
    assert os.path.exists(tmp_path / "iter_5.pth")


# ---------------------------------------------------------------------------- vectorised bf16 elementwise kernels, wide channels
@pytest.mark.parametrize("C", [64, 800, 1152, 2304, 4608])
def test_act_backward_bf16_wide_channels(C):
    """da_conv_act_backward on bf16 activations takes the 16-byte kernel for every C % 8 == 0: G = C/8 <= 128 column groups
    (rows split over the block, idle threads when 256 % G != 0) and G > 128 (a thread walks several groups)."""
    import ctypes
    g = torch.Generator(device=DEV).manual_seed(C)
    M = 333
    dy = torch.randn(M, C, device=DEV, generator=g).to(torch.bfloat16)
    y = torch.randn(M, C, device=DEV, generator=g).to(torch.bfloat16)
    scale = torch.rand(C, device=DEV, generator=g) + 0.5
    dz = torch.empty_like(dy)
    dshift = torch.empty(C, device=DEV)
    dvdot = torch.empty(C, device=DEV)
    desc = F_._conv_desc(M, 1, 1, 8, C, 1, 1, 1, 0, "umma_bf16", torch.bfloat16, torch.bfloat16)
    ws = F_.workspace(uda._lib.lib.da_conv_workspace_bytes(ctypes.byref(desc)), torch.device(DEV), "conv")
    uda._lib.check(uda._lib.lib.da_conv_act_backward(ctypes.byref(desc), F_._ptr(dy), F_._ptr(y), F_._ptr(scale), 1, 0.0, 0, F_._ptr(dz),
                                                     F_._ptr(dshift), F_._ptr(dvdot), F_._ptr(ws), ws.numel(), None), "act_backward")
    torch.cuda.synchronize()
    d = dy.float() * (y.float() > 0)
    assert torch.equal(dz, (d * scale).to(torch.bfloat16))
    assert rel_err(dshift, d.sum(0)) <= 1e-5
    assert rel_err(dvdot, (d * y.float()).sum(0)) <= 1e-5


@pytest.mark.parametrize("C", [64, 1152, 4608])
def test_global_avgpool_bf16_vectorised(C):
    g = torch.Generator(device=DEV).manual_seed(C + 1)
    x = torch.randn(2, 9, 13, C, device=DEV, generator=g).to(torch.bfloat16).requires_grad_(True)
    yv = F_.global_avgpool(x)
    ref = x.detach().float().mean(dim=(1, 2))
    assert rel_err(yv.float().view(2, C), ref) <= 1e-5
    cot = torch.randn(yv.shape, device=DEV, generator=g).to(yv.dtype)
    (dx,) = torch.autograd.grad(yv, x, cot)
    exp = (cot.float().view(2, 1, 1, C) / (9 * 13)).expand(2, 9, 13, C).to(torch.bfloat16)
    assert torch.equal(dx, exp)


# ---------------------------------------------------------------------------- the benchmarked configuration, at full size
def _bench_shape_inputs():
    """bench.py's workload: one source+target pair, C5 [2,2048,64,128], 512 RoIs per image with log-uniform sizes."""
    g = torch.Generator().manual_seed(123)
    c5 = torch.relu(torch.randn(2, 64, 128, 2048, generator=g))            # NHWC storage, as bench.make_host_inputs
    u = torch.rand(2, 512, 4, generator=g)
    x1, y1 = u[..., 0] * (128 * 16 - 33), u[..., 1] * (64 * 16 - 33)
    lo, hi = torch.log(torch.tensor(16.0)), torch.log(torch.tensor(512.0))
    w, h = torch.exp(lo + u[..., 2] * (hi - lo)), torch.exp(lo + u[..., 3] * (hi - lo))
    boxes = torch.stack([x1, y1, torch.clamp(x1 + w, max=2048.0), torch.clamp(y1 + h, max=1024.0)], -1).contiguous()
    return c5, boxes


@pytest.mark.parametrize("engine,roi_layout", [("umma_bf16", "rchw"), ("umma_bf16x6", "rchw"), ("umma_bf16", "rhwc")])
def test_daf_org_hot_path_at_bench_shape_vs_oracle(engine, roi_layout):
    """The configuration bench.py times (hotpath.DAFOrgHotPath, C5 [2,2048,64,128], 2x512 RoIs, shared FC 100352 -> 1024 ->
    1024, InstanceAlignmentHead over the 1024 RoIs, L1 + L4 + L7, full backward) against the CPU oracle AT FULL SIZE:
    fp64 torch for heads / FCs / losses, oracle/roi_align_ref.c for RoIAlign forward and its transposed map.  Dropout off
    (eval) so both sides see the same function.  What is compared:
      * the three losses of the dict                                    (all RoIs, all pixels)
      * d loss / d C5                                                    (whole tensor, max-norm relative)
      * d loss / d FC1.weight on 8 output rows x all 100352 columns      (slice: the oracle forms dz^T X for those rows only)
      * d loss / d of every instance-head weight                         (whole tensors)
    umma_bf16x6 (tcgen05, fp32-class) is held to the fp32 bar; umma_bf16 is compared with the oracle evaluated at the same
    bf16 storage points (q='bf16') and stated separately.
    ReLU subgradients: the path has ~11 M ReLU units; a unit whose pre-activation is within rounding of 0 is switched on by
    one side and off by the other, and ONE such unit of FC1 moves the gradient of a whole RoI by ~1/sqrt(#active units)
    (measured without this provision: fp32-class engine, ~40 such units, dC5 max-norm error 1.5e-2 with all losses <= 4e-6).
    The oracle therefore takes the CUDA path's on/off decision inside a band |z| <= delta*max|z| (delta = 4e-5 for the
    fp32-class engine, 1e-2 for bf16) and its own decision outside; the test asserts that NO disagreement lies outside the band."""
    from unsupervised_domain_adaptation_object_detection_implementation_b200 import hotpath
    from helpers import ReluLikeCuda
    import torch.nn.functional as tF
    uda.set_engine(engine)
    bf16 = engine == "umma_bf16"
    q = "bf16" if bf16 else None
    torch.manual_seed(0)
    model = hotpath.DAFOrgHotPath(2048, 16, 1024).eval()
    seeded.fill_state_(model, 7, "bench.")                                  # O(1) activations everywhere (Q17)
    if roi_layout == "rhwc":
        # the SAME model through the reference-order state_dict: the first shared FC then holds its columns in (bin, channel)
        # order, state_dict() below gives them back in the reference's
        sd0 = model.state_dict()
        model = hotpath.DAFOrgHotPath(2048, 16, 1024, roi_layout="rhwc").eval()
        model.load_state_dict(sd0)
        del sd0
    c5_nhwc, boxes = _bench_shape_inputs()
    if bf16:
        c5_nhwc = c5_nhwc.bfloat16().float()
    rois = torch.cat([torch.cat([torch.full((512, 1), float(i)), boxes[i]], 1) for i in range(2)])
    labels = (torch.arange(1024) >= 512).long()
    gt = torch.tensor([0, 1])
    sd = {k: v.detach().double() for k, v in model.state_dict().items()}

    # ---- CUDA path (first: the oracle needs its ReLU decisions)
    model = model.to(DEV)
    acts, orig = [], F_.dense_layer

    def recording_dense_layer(*a, **k):
        y = orig(*a, **k)
        if k.get("relu"):
            acts.append(y.detach())
        return y

    chain_orig = F_.instance_head_chain

    def recording_chain(*a, **k):         # bf16 engine: the instance head is ONE kernel; its stored activations h1, h2
        loss_c, pred_c = chain_orig(*a, **k)
        sv = loss_c.grad_fn.keep[2]
        if k.get("pre") is not None:      # the last shared FC ran inside the kernel: its output is the kernel's stored x
            acts.append(loss_c.grad_fn.keep[0].detach())
        acts.extend([sv["h1"].detach(), sv["h2"].detach()])
        return loss_c, pred_c

    tail_orig = F_.conv_pixel_loss

    def recording_tail(*a, **k):          # H1: conv1 + fused tail; its stored conv1 activation
        loss_t, logits_t = tail_orig(*a, **k)
        acts.insert(0, loss_t.grad_fn.saved_tensors[2].detach())
        return loss_t, logits_t

    x = c5_nhwc.to(DEV).to(F_.act_dtype()).permute(0, 3, 1, 2).requires_grad_(True)
    F_.dense_layer, F_.instance_head_chain, F_.conv_pixel_loss = recording_dense_layer, recording_chain, recording_tail
    try:
        losses = model.forward_train(x, [boxes[0].to(DEV), boxes[1].to(DEV)], [0, 1])
    finally:
        F_.dense_layer, F_.instance_head_chain, F_.conv_pixel_loss = orig, chain_orig, tail_orig
    total, _ = hotpath.parse_losses(losses)
    total.backward()
    assert len(acts) == 5                                                   # H1 conv1, shared FC1, FC2, instance fc1, fc2
    with torch.no_grad():
        img_feat_cuda = model.da_head_top(x.detach())
    cuda_acts = [acts[0].permute(0, 3, 1, 2).float().cpu(), img_feat_cuda.float().cpu()] + \
        [a.reshape(1024, -1).float().cpu() for a in acts[1:]]

    # ---- oracle (CPU)
    vtol = BF16_TOL if bf16 else FP32_TOL
    relu = ReluLikeCuda(cuda_acts, 1e-2 if bf16 else 4e-5)        # band: a few x the rounding noise of a pre-activation
    torch.set_num_threads(os.cpu_count() or 8)
    Q = (lambda t: da_oracle._q(t, q))
    sub = lambda p: {k[len(p):]: v for k, v in sd.items() if k.startswith(p)}
    c5r = c5_nhwc.permute(0, 3, 1, 2).double().requires_grad_(True)
    img_feat = da_oracle.img_alignment_head(c5r, sub("da_head_top."), q=q, relu=relu)
    pooled_np, grid_ref, _ = oracle_roi.roi_align_forward(c5_nhwc.permute(0, 3, 1, 2).contiguous().numpy(), rois.numpy(), 7, 1 / 16,
                                                          threads=os.cpu_count() or 8)
    pooled = torch.from_numpy(pooled_np).double().requires_grad_(True)
    z1 = tF.linear(Q(pooled.flatten(1)), Q(sd["bbox_head.shared_fcs.0.weight"]), sd["bbox_head.shared_fcs.0.bias"])
    z1.retain_grad()
    f1 = Q(relu(z1))
    f2 = Q(relu(tF.linear(f1, Q(sd["bbox_head.shared_fcs.1.weight"]), sd["bbox_head.shared_fcs.1.bias"])))
    sd_ins = {k: v.requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sub("local_da.").items()}
    pred = torch.sigmoid(da_oracle.instance_alignment_logits(f2, sd_ins, q=q, relu=relu))
    ref = dict(globle_da_loss=0.1 * da_oracle.daf_image_loss(img_feat, gt), local_da_loss=0.1 * da_oracle.ce2(pred, labels),
               consistency_loss=0.1 * da_oracle.consistency_loss(img_feat, pred, labels))
    sum(ref.values()).backward()
    c5_grad_ref = c5r.grad + torch.from_numpy(oracle_roi.roi_align_backward(pooled.grad.float().numpy(), rois.numpy(),
                                                                            (2, 2048, 64, 128), 7, 1 / 16))
    rows = torch.tensor([0, 1, 127, 128, 511, 640, 1000, 1023])
    dw1_ref = z1.grad[:, rows].t() @ Q(pooled.detach().flatten(1))          # [8, 100352]
    # L7 is |m - s_r|: well-posed for a gradient comparison only while no RoI sits at the kink
    margin = (torch.sigmoid(img_feat.detach()).mean() - torch.sigmoid(pred.detach()[torch.arange(1024), labels])).abs()
    assert float(margin.min()) > (1e-2 if bf16 else 1e-4)

    # ---- compare
    ltol, gtol, wtol = (BF16_TOL, 2 * BF16_TOL, 5e-2) if bf16 else (FP32_TOL, 5e-5, 5e-5)
    gx = x.grad.double().cpu()
    e_max, e_fro = rel_err(gx, c5_grad_ref), float((gx - c5_grad_ref).norm() / c5_grad_ref.norm())
    print(f"\n[bench-shape {engine}] losses " + ", ".join(f"{k}: {abs(float(losses[k]) - float(ref[k])) / abs(float(ref[k])):.2e}" for k in ref) +
          f" | dC5 max-norm {e_max:.3e} frobenius {e_fro:.3e} | ReLU units near 0: {relu.near}, decided by the CUDA path: {relu.flips}, "
          f"disagreements outside the band: {relu.outside}")
    assert relu.outside == 0
    for k in ref:
        assert abs(float(losses[k]) - float(ref[k])) <= ltol * abs(float(ref[k])), (k, float(losses[k]), float(ref[k]))
    assert e_max <= gtol
    dw1 = model.bbox_head.to_reference_order(model.bbox_head.shared_fcs[0].weight.grad[rows.to(DEV)])
    assert float((dw1.double().cpu() - dw1_ref).norm() / dw1_ref.norm()) <= wtol
    for k, v in sd_ins.items():
        if v.grad is not None and v.dim() >= 2:
            got = dict(model.local_da.named_parameters())[k].grad
            assert got is not None, k
            e = float((got.double().cpu() - v.grad).norm() / v.grad.norm().clamp_min(1e-30))
            assert e <= wtol, (k, e)


def test_roi_align_backward_reuses_forward_preparation_only_when_valid():
    """da_roi_align_backward_prepared: the backward skips the preparation launch when the shared workspace still holds the
    forward's tap tables for the SAME RoI set (functional._ROI_PREP token); any roi_align call in between (other RoIs, other
    map) invalidates the token and the backward prepares again.  All three routes give bit-identical gradients, for the
    tensor-core (bf16) and the CUDA-core (fp32) kernels."""
    for dtype in (torch.bfloat16, torch.float32):
        feat = seeded.feature_map("prep.f", (2, 128, 24, 40), 0).to(DEV).to(dtype).contiguous(memory_format=torch.channels_last)
        rois = seeded.synthetic_rois(48, 2, 24 * 16, 40 * 16, 1).to(DEV)
        other = seeded.synthetic_rois(16, 2, 24 * 16, 40 * 16, 2).to(DEV)
        cot = seeded.seeded_tensor("prep.c", (96, 128, 7, 7), 0).to(DEV).to(dtype)
        grads, routes = [], []
        for interleave in (False, True, False):
            x = feat.clone(memory_format=torch.preserve_format).requires_grad_(True)
            out = F_.roi_align(x, rois, 7, 1 / 16)
            if interleave:
                F_.roi_align(feat, other, 7, 1 / 16)          # overwrites the workspace: the token must not match any more
            tok = F_._ROI_PREP.get(x.device.index)
            routes.append(tok is not None and tok[1] == out.grad_fn.saved_tensors[0].data_ptr())   # what the backward will find
            (g,) = torch.autograd.grad(out, x, cot)
            grads.append(g.clone())
        assert routes == [True, False, True]
        assert torch.equal(grads[0], grads[1]) and torch.equal(grads[0], grads[2])
        assert float(grads[0].float().abs().max()) > 0


@pytest.mark.parametrize("layout", ["rchw", "rhwc"])
def test_roi_align_tensor_core_backward_at_bench_size_vs_oracle(layout):
    """(layout = memory order of the RoI tensor: the reference's [R,C,7,7] or bin-major [R,7,7,C].)
    bf16 tcgen05 RoIAlign backward at the benchmarked size (C=2048, 64x128, 1024 RoIs over 2 images): the whole
    [2,2048,64,128] gradient against the fp64 oracle backward on the same bf16-rounded cotangent.  Tolerance: the
    tap weights are rounded to bf16 (2^-9 each, independent across the taps that are summed) and the result is rounded
    once to bf16 -> 1e-2 of the max magnitude, as for the small shapes above."""
    N, C, H, W, R = 2, 2048, 64, 128, 1024
    c5, boxes = _bench_shape_inputs()
    rois = torch.cat([torch.cat([torch.full((512, 1), float(i)), boxes[i]], 1) for i in range(2)])
    f = c5.to(DEV).to(torch.bfloat16).permute(0, 3, 1, 2).requires_grad_(True)
    out = F_.roi_align(f, rois.to(DEV), 7, 1 / 16, out_layout=layout)
    g = torch.Generator().manual_seed(5)
    cot = torch.randn(R, C, 7, 7, generator=g).to(torch.bfloat16)
    (g_tc,) = torch.autograd.grad(out, f, cot.to(DEV) if layout == "rchw" else cot.to(DEV).permute(0, 2, 3, 1).contiguous())
    assert g_tc.dtype == torch.bfloat16
    if layout == "rhwc":
        out = out.permute(0, 3, 1, 2)
    gref = torch.from_numpy(oracle_roi.roi_align_backward(cot.float().numpy(), rois.numpy(), (N, C, H, W), 7, 1 / 16))
    assert rel_err(g_tc.float(), gref) <= 1e-2
    # forward at the same size against the oracle on the bf16-rounded map (all channels, all RoIs)
    ref, grid_ref, _ = oracle_roi.roi_align_forward(f.detach().float().cpu().numpy(), rois.numpy(), 7, 1 / 16, threads=os.cpu_count() or 8)
    assert rel_err(out.float(), torch.from_numpy(ref)) <= 1e-2


# ---------------------------------------------------------------------------- R3 StandardRoIHeadDA_v5: value parity
@pytest.mark.parametrize("engine,tol", [("simt_f32", FP32_TOL), ("umma_bf16x6", FP32_TOL)])
def test_roi_head_da_v5_values_vs_oracle(engine, tol):
    """StandardRoIHeadDA_v5.forward_train (standard_roi_head_da_v5.py:162-227) with the RANDOM parts pinned (assign + sample are
    replaced by fixed per-image samples): returned losses (source image only), bbox_feats = [src, tar], bbox_cls = [src, tar] and the
    gradients of the bbox-head parameters against the oracle transcription (da_oracle.roi_head_da_v5 over the RoIAlign oracle)."""
    from unsupervised_domain_adaptation_object_detection_implementation_b200 import detection
    uda.set_engine(engine)
    C, ncls, stride, H, W = 64, 3, 16, 12, 20
    head = detection.StandardRoIHeadDA_v5(
        bbox_roi_extractor=dict(type="SingleRoIExtractor", roi_layer=dict(type="RoIAlign", output_size=7, sampling_ratio=0),
                                out_channels=C, featmap_strides=[stride]),
        bbox_head=dict(type="Shared2FCBBoxHead", in_channels=C, fc_out_channels=128, roi_feat_size=7, num_classes=ncls,
                       reg_class_agnostic=False))
    seeded.fill_state_(head, 3, "r3.")
    x = seeded.feature_map("r3.x", (2, C, H, W), 3)
    sampled = []
    for i, n in enumerate((24, 17)):
        boxes = seeded.synthetic_rois(n, 1, H * stride, W * stride, 10 + i, 12.0, 150.0)[:, 1:]
        labels = (seeded.seeded_tensor(f"r3.lab{i}", (n,), 3, "uniform") * 3.5 - 1.6).clamp(0, ncls).long()
        targets = seeded.seeded_tensor(f"r3.tgt{i}", (n, 4), 3, scale=0.5)
        sampled.append((boxes, labels, targets, labels < ncls))
    # oracle
    sd = {k: v.detach().double().requires_grad_(True) for k, v in head.bbox_head.state_dict().items()}
    pooled = [torch.from_numpy(oracle_roi.roi_align_forward(x.numpy(), torch.cat([torch.full((len(s[0]), 1), float(i)), s[0]], 1).numpy(),
                                                            7, 1 / stride)[0]).double() for i, s in enumerate(sampled)]
    ref_losses, ref_feats, ref_cls = da_oracle.roi_head_da_v5(pooled, sampled, sd, ncls, [0, 1])
    cot = [seeded.seeded_tensor(f"r3.cot{i}", tuple(f.shape), 3) for i, f in enumerate(ref_feats)]
    (ref_losses["loss_cls"] + ref_losses["loss_bbox"] + sum((f * c.double()).sum() for f, c in zip(ref_feats, cot))).backward()
    # CUDA path: the sampler is pinned to the same samples
    head = head.to(DEV)
    it = iter([tuple(t.to(DEV) for t in s) for s in sampled])
    head._sample = lambda *a, **k: next(it)
    metas = [dict(img_shape=(H * stride, W * stride, 3)) for _ in range(2)]
    losses, feats, cls = head.forward_train([x.to(DEV)], metas, [None, None], [None, None], [None, None], gt_da=[0, 1])
    assert set(losses) == {"loss_cls", "loss_bbox", "acc"} and len(feats) == 2 and len(cls) == 2
    for k in ("loss_cls", "loss_bbox"):
        assert abs(float(losses[k]) - float(ref_losses[k])) <= tol * abs(float(ref_losses[k])), k
    assert abs(float(losses["acc"]) - float(ref_losses["acc"])) <= 1e-3
    for a, b in zip(feats, ref_feats):
        assert a.shape == b.shape and rel_err(a.float(), b) <= tol
    for a, b in zip(cls, ref_cls):
        assert rel_err(a.float(), b) <= tol
    (losses["loss_cls"] + losses["loss_bbox"] + sum((f.float() * c.to(DEV)).sum() for f, c in zip(feats, cot))).backward()
    for k, p in head.bbox_head.named_parameters():
        assert p.grad is not None, k
        assert rel_err(p.grad, sd[k].grad) <= 5 * tol, k


# ---------------------------------------------------------------------------- instance head + CE as one kernel (csrc/chain.cu)
def _chain_case(name, R, seed=0):
    from unsupervised_domain_adaptation_object_detection_implementation_b200 import da_heads
    m = build_head(name, seed)
    x = seeded.feature_map(f"chain.{name}.x", (R, 1024), seed).bfloat16().float()
    labels = (torch.arange(R) >= R // 2).long()
    gpred = seeded.seeded_tensor(f"chain.{name}.gp", (R, 2), seed, scale=0.3 / R)
    return m, x, labels, gpred


@pytest.mark.parametrize("name", ["instance_alignment", "instance_alignment_daf"])
@pytest.mark.parametrize("R", [24, 300, 1024])
def test_instance_head_chain_vs_oracle_and_layerwise(name, R):
    """da_instance_fc_forward/backward (ONE persistent kernel each: NonLocalBlock projections, query-axis softmax, attention,
    FC stack, sigmoid + CE; data / weight / bias gradients with the GRL folded in) against
      (a) the oracle evaluated at the same bf16 storage points (q='bf16'), loss + pred + dx + every parameter gradient,
      (b) the layer-by-layer CUDA path (same kernels as the goldens pin), eval mode.
    R = 24 (golden size, one partial tile), 300 (ragged: rows, attention columns and the K of P*g all have tails), 1024 (bench)."""
    from unsupervised_domain_adaptation_object_detection_implementation_b200 import da_heads
    uda.set_engine("umma_bf16")
    m, x, labels, gpred = _chain_case(name, R)
    # (a) oracle
    sd = {k: v.detach().double().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in m.state_dict().items()}
    xr = x.double().requires_grad_(True)
    logits = (da_oracle.instance_alignment_logits if name == "instance_alignment" else da_oracle.instance_alignment_daf_logits)(xr, sd, q="bf16")
    pred_r = torch.sigmoid(logits)
    loss_r = da_oracle.ce2(pred_r, labels)
    (0.1 * loss_r + (pred_r * gpred.double()).sum()).backward()
    # chain
    m = m.to(DEV)
    xc = x.to(DEV).requires_grad_(True)
    loss, pred = m.forward_loss(xc, labels.to(DEV))
    assert pred.shape == (R, 2) and loss.dim() == 0
    (0.1 * loss + (pred * gpred.to(DEV)).sum()).backward()
    assert abs(float(loss) - float(loss_r)) <= BF16_TOL * abs(float(loss_r))
    assert rel_err(pred, pred_r) <= BF16_TOL
    # every gradient of the backward chain (dz, dz2, dz1, dt, dY, dS, dtheta|dphi|dg) is stored in bf16: bounded in the
    # Frobenius norm (3e-2) and, looser, element-wise against the max magnitude
    e_fro = float((xc.grad.double().cpu() - xr.grad).norm() / xr.grad.norm())
    print(f"\n[chain {name} R={R}] loss {abs(float(loss) - float(loss_r)) / abs(float(loss_r)):.2e} pred {rel_err(pred, pred_r):.2e} "
          f"dx max-norm {rel_err(xc.grad, xr.grad):.2e} frobenius {e_fro:.2e}")
    chain_grads, errs = {}, {}
    for k, p in m.named_parameters():
        if sd[k].grad is None:
            assert p.grad is None, k
            continue
        assert p.grad is not None and p.grad.shape == p.shape, k
        a, b = p.grad.double().cpu(), sd[k].grad
        errs[k] = float((a - b).norm() / b.norm().clamp_min(1e-30))
        chain_grads[k] = p.grad.clone()
        p.grad = None
    print("   param grad frobenius errors:", {k: f"{v:.1e}" for k, v in errs.items()})
    assert e_fro <= BF16_TOL and rel_err(xc.grad, xr.grad) <= 0.12
    assert all(v <= 5e-2 for v in errs.values()), errs
    # (b) layer-by-layer path
    da_heads.USE_CHAIN = False
    try:
        xl = x.to(DEV).requires_grad_(True)
        loss_l, pred_l = m.forward_loss(xl, labels.to(DEV))
        (0.1 * loss_l + (pred_l * gpred.to(DEV)).sum()).backward()
    finally:
        da_heads.USE_CHAIN = True
    # two bf16 paths with different rounding points (the layer-wise path rounds the NonLocalBlock output twice, keeps fc3 and
    # dz in fp32, ...): a sanity bound in the Frobenius norm, the parity statement is (a)
    fro = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
    assert abs(float(loss) - float(loss_l)) <= 2e-3 * abs(float(loss_l)) and rel_err(pred, pred_l) <= 2e-2
    assert fro(xc.grad, xl.grad) <= 0.1
    for k, p in m.named_parameters():
        if k in chain_grads:
            assert fro(chain_grads[k], p.grad) <= 0.12, k


@pytest.mark.parametrize("name", ["instance_alignment", "instance_alignment_daf"])
@pytest.mark.parametrize("R", [300, 1024])
def test_instance_head_chain_feeding_layer_matches_separate_layers(name, R):
    """desc.C0 > 0: the last shared FC of the bbox head (convfc_bbox_head.py:229-237) runs INSIDE the chain kernel, forward and
    backward, and the kernel applies the ReLU derivative of the FC before it and returns that layer's bias gradient
    (hotpath.SharedFCs.split + dense_layer(preact_grad=True)).  Same storage points (bf16 activations) as the separate
    launches, so loss, pred, the gradient into the RoI features and EVERY parameter gradient of both FCs and of the head agree
    to accumulation-order noise."""
    from unsupervised_domain_adaptation_object_detection_implementation_b200 import da_heads, hotpath
    uda.set_engine("umma_bf16")
    torch.manual_seed(0)
    fcs = hotpath.SharedFCs(64, 2, 1024).to(DEV)
    seeded.fill_state_(fcs, 3, "feed.")
    with torch.no_grad():
        for fc in fcs.shared_fcs:
            fc.bias.add_(0.05)            # non-trivial biases (the module initialises them to 0)
    head = build_head(name, 1).to(DEV).eval()
    roi = seeded.feature_map("feed.roi", (R, 64, 2, 2), 0).to(DEV).to(torch.bfloat16)
    labels = (torch.arange(R, device=DEV) >= R // 2).long()
    gpred = seeded.seeded_tensor("feed.gp", (R, 2), 0, scale=0.3 / R).to(DEV)
    outs = []
    for feed in (True, False):
        x = roi.clone().requires_grad_(True)
        if feed:
            assert fcs.can_feed_chain(x)
            xin, pre = fcs.split(x)
            loss, pred = head.forward_loss(None, labels, pre=(xin,) + pre)
        else:
            loss, pred = head.forward_loss(fcs(x), labels)
        (0.1 * loss + (pred * gpred).sum()).backward()
        grads = {"x": x.grad.float().clone()}
        for mod, pfx in ((fcs, "fcs."), (head, "head.")):
            for k, p in mod.named_parameters():
                if p.grad is not None:
                    grads[pfx + k] = p.grad.clone()
                    p.grad = None
        outs.append((float(loss), pred.detach().clone(), grads))
    (la, pa, ga), (lb, pb, gb) = outs
    fro = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
    assert abs(la - lb) <= 1e-4 * abs(lb) and rel_err(pa, pb) <= 1e-3
    assert set(ga) == set(gb), set(ga) ^ set(gb)
    errs = {k: fro(ga[k], gb[k]) for k in ga}
    print(f"\n[chain feed {name} R={R}]", {k: f"{v:.1e}" for k, v in errs.items()})
    assert all(v <= 2e-2 for v in errs.values()), errs
    assert "fcs.shared_fcs.0.bias" in ga and "fcs.shared_fcs.1.weight" in ga


def test_instance_head_chain_dropout_matches_layerwise_masks():
    """Training mode: the chain draws its two dropout seeds in the same order as the layer-by-layer path and hashes the same
    element index (row * width + column), so both paths drop the SAME units; the stored activations double as the ReLU/dropout
    derivative mask in the chain's backward."""
    from unsupervised_domain_adaptation_object_detection_implementation_b200 import da_heads
    uda.set_engine("umma_bf16")
    m, x, labels, gpred = _chain_case("instance_alignment", 256, seed=2)
    m = m.to(DEV).train()
    outs = []
    for use in (True, False):
        da_heads.USE_CHAIN = use
        try:
            torch.manual_seed(1234)
            xc = x.to(DEV).requires_grad_(True)
            loss, pred = m.forward_loss(xc, labels.to(DEV))
            (loss + (pred * gpred.to(DEV)).sum()).backward()
            outs.append((float(loss), pred.detach().clone(), xc.grad.clone(), {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}))
            for p in m.parameters():
                p.grad = None
        finally:
            da_heads.USE_CHAIN = True
    (la, pa, ga, wa), (lb, pb, gb, wb) = outs
    fro = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
    assert abs(la - lb) <= 3e-3 * abs(lb) and rel_err(pa, pb) <= 2e-2
    assert fro(ga, gb) <= 0.1               # different masks would give O(1)
    assert set(wa) == set(wb)
    for k in wa:
        assert fro(wa[k], wb[k]) <= 0.12, k
    torch.manual_seed(99)                                             # another seed: other masks
    xc = x.to(DEV)
    _, p2 = m.forward_loss(xc.requires_grad_(True), labels.to(DEV))
    assert float((p2 - pa).abs().max()) > 0


def test_instance_head_chain_golden_and_packed_projection_state(golden):
    """pred of the chain against the reference class's golden output; the packed theta|phi|g storage keeps parameter names,
    shapes and values, survives load_state_dict and is re-packed after .to()."""
    uda.set_engine("umma_bf16")
    g = golden("head_instance_alignment.pt")
    m = build_head("instance_alignment", g["seed"]).to(DEV)
    before = {k: v.clone() for k, v in m.state_dict().items()}
    x = g["x"].to(DEV).requires_grad_(True)
    _, pred = m.forward_loss(x, torch.zeros(x.shape[0], dtype=torch.long, device=DEV))
    assert rel_err(pred, g["out"][0]) <= BF16_TOL
    nlb = m.nlb
    ws = [nlb.conv_theta.weight, nlb.conv_phi.weight, nlb.conv_g.weight]
    n = ws[0].numel() * 4
    assert ws[1].data_ptr() == ws[0].data_ptr() + n and ws[2].data_ptr() == ws[0].data_ptr() + 2 * n      # one buffer
    after = m.state_dict()
    assert set(after) == set(before) and all(torch.equal(after[k], before[k]) for k in before)
    m.load_state_dict({k: v * 0.5 if "conv_phi" in k else v for k, v in before.items()})
    assert ws[1].data_ptr() == ws[0].data_ptr() + n                                                       # in-place load keeps it
    _, pred2 = m.forward_loss(x, torch.zeros(x.shape[0], dtype=torch.long, device=DEV))
    assert float((pred2 - pred).abs().max()) > 0                                                          # operand copy refreshed
    m2 = m.float().cpu().to(DEV)                                                                          # storages separated
    _, pred3 = m2.forward_loss(x, torch.zeros(x.shape[0], dtype=torch.long, device=DEV))
    assert rel_err(pred3, pred2) <= 1e-6


# ---------------------------------------------------------------------------- fused pixel tail (north_star kernel 1) + BCE / focal modes
def _tail_reference(mode, p, gt, gamma=2.0, alpha=0.25):
    """Per-pixel loss on logits p [N,1,H,W] (double) against the image's domain label."""
    if mode == "daf_sq_batch":
        return da_oracle.daf_image_loss(p, gt)
    if mode == "daf_sq_image":
        return da_oracle.patch_loss(p, gt)
    t = gt.to(p.dtype).view(-1, 1, 1, 1).expand_as(p)
    if mode == "bce":
        return torch.nn.functional.binary_cross_entropy_with_logits(p, t)                  # mean over all pixels
    # py_sigmoid_focal_loss (mmdet/models/losses/focal_loss.py:12-57), reduction='mean'
    s = p.sigmoid()
    pt = (1 - s) * t + s * (1 - t)
    fw = (alpha * t + (1 - alpha) * (1 - t)) * pt.pow(gamma)
    return (torch.nn.functional.binary_cross_entropy_with_logits(p, t, reduction="none") * fw).mean()


@pytest.mark.parametrize("engine,tol", [("simt_f32", FP32_TOL), ("umma_bf16x6", FP32_TOL), ("umma_bf16", BF16_TOL)])
@pytest.mark.parametrize("mode", ["daf_sq_batch", "daf_sq_image", "bce", "focal"])
@pytest.mark.parametrize("head", ["img_alignment", "local_alignment"])
def test_pixel_head_fused_loss_modes_vs_oracle(golden, head, mode, engine, tol):
    """da_grl_conv_loss_forward/backward: the pixel-level heads with their loss fused into the tail kernel, all four per-pixel
    modes (the reference's L1 / L2 and the plain sigmoid-BCE / focal modes, against F.binary_cross_entropy_with_logits and the
    py_sigmoid_focal_loss formula), loss + logits + input gradient (GRL folded) + every parameter gradient, plus a second
    consumer of the logits (as the consistency loss is)."""
    uda.set_engine(engine)
    g = golden(f"head_{head}.pt")
    m = build_head(head, g["seed"])
    q = "bf16" if engine == "umma_bf16" else None
    x = g["x"].bfloat16().float() if q else g["x"]
    gt = torch.tensor([0, 1])
    cot = g["cot"][0] * 0.05
    sd = {k: v.detach().double().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in m.state_dict().items()}
    xr = x.double().requires_grad_(True)
    pr = HEADS[head][1](xr, sd, q=q)
    lr_ = _tail_reference(mode, pr, gt)
    (0.3 * lr_ + (pr * cot.double()).sum()).backward()
    m = m.to(DEV)
    xc = x.to(DEV).requires_grad_(True)
    loss, feat = m.forward_loss(xc, gt.to(DEV), mode=mode)
    assert feat.shape == pr.shape
    assert rel_err(feat, pr) <= tol
    assert abs(float(loss) - float(lr_)) <= tol * abs(float(lr_)), (float(loss), float(lr_))
    (0.3 * loss + (feat * cot.to(DEV)).sum()).backward()
    gtol = 2 * tol if q else 5 * tol
    assert rel_err(xc.grad, xr.grad) <= gtol
    for k, p in m.named_parameters():
        if sd[k].grad is None:
            continue
        assert p.grad is not None, k
        a, b = p.grad.double().cpu(), sd[k].grad
        assert float((a - b).norm() / b.norm().clamp_min(1e-30)) <= (5e-2 if q else 5 * tol), k


def test_pixel_tail_matches_unfused_path_at_bench_size():
    """H1 + L1 at the benchmarked size (C5 [2,2048,64,128]): fused tail vs the separate pixel_head / pixel_domain_loss kernels
    (same GEMM in front): identical logits, loss to 1e-6, gradients to bf16 rounding of dz."""
    uda.set_engine("umma_bf16")
    from unsupervised_domain_adaptation_object_detection_implementation_b200 import da_heads
    torch.manual_seed(0)
    m = da_heads.ImgAlignmentHead(2048)
    seeded.fill_state_(m, 5, "tailbench.")
    m = m.to(DEV)
    g = torch.Generator(device=DEV).manual_seed(3)
    x = torch.relu(torch.randn(2, 64, 128, 2048, device=DEV, generator=g)).to(torch.bfloat16).permute(0, 3, 1, 2)
    gt = torch.tensor([0, 1], device=DEV)
    xa = x.detach().requires_grad_(True)
    la, fa = m.forward_loss(xa, gt)
    la.backward()
    ga = {k: p.grad.clone() for k, p in m.named_parameters()}
    for p in m.parameters():
        p.grad = None
    xb = x.detach().requires_grad_(True)
    fb = m(xb)
    lb = da_losses.daf_image_loss(fb, gt)
    lb.backward()
    assert torch.equal(fa, fb) or rel_err(fa, fb) <= 1e-6
    assert abs(float(la) - float(lb)) <= 1e-6 * abs(float(lb))
    assert rel_err(xa.grad.float(), xb.grad.float()) <= 2e-2
    for k, p in m.named_parameters():
        assert float((ga[k] - p.grad).norm() / p.grad.norm().clamp_min(1e-30)) <= 2e-2, k


# ---------------------------------------------------------------------------- hot-path modules beyond DAF-Org
def test_fpn_hot_path_runs_under_cuda_graph_and_matches_eager():
    """FPNHotPath (BASELINE config 2b: heads + multi-level RoIAlign + shared FCs + instance head on P2..P5): the multi-level
    extractor partitions the RoIs on the device (no nonzero / host sync), so the whole forward+backward is CUDA-graph
    capturable; the replay reproduces the eager losses and input gradients bit for bit (eval: no dropout)."""
    from unsupervised_domain_adaptation_object_detection_implementation_b200 import hotpath
    uda.set_engine("umma_bf16")
    torch.manual_seed(0)
    m = hotpath.FPNHotPath(64, (4, 8, 16, 32), 1024).to(DEV).eval()        # the instance head is 1024 wide (instance_da.py:50)
    seeded.fill_state_(m, 1, "fpn.")
    feats = [seeded.feature_map(f"fpn.f{i}", (2, 64, 64 >> i, 96 >> i), 1).to(DEV).to(torch.bfloat16) for i in range(4)]
    props = [seeded.synthetic_rois(40, 1, 256, 384, s, 8.0, 300.0)[:, 1:].to(DEV) for s in (0, 1)]

    def step(_unused=None):
        xs = [f.detach().clone().requires_grad_(True) for f in feats]     # leaves live on the stream that runs the step (as in bench.py)
        losses = m.forward_train(xs, props, [0, 1])
        total, log = hotpath.parse_losses(losses)
        grads = torch.autograd.grad(total, xs)
        return total.detach(), [g.detach() for g in grads]

    xs = None
    for _ in range(2):
        t_eager, g_eager = step(xs)
    assert torch.isfinite(t_eager) and all(torch.isfinite(g.float()).all() and float(g.float().abs().sum()) > 0 for g in g_eager)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step(xs)
    torch.cuda.current_stream().wait_stream(side)
    gph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gph):
        t_cap, g_cap = step(xs)
    gph.replay()
    torch.cuda.synchronize()
    assert float(t_cap) == float(t_eager)
    for a, b in zip(g_cap, g_eager):
        assert torch.equal(a, b)
    losses = m.forward_train(feats, props, [0, 1])
    assert set(losses) == {"local_da_loss", "consistency_loss", "globle_da_loss"} and len(losses["globle_da_loss"]) == 4


def test_deep_hot_path_losses_vs_oracle():
    """DeepHotPath (DAFasterRCNN_Deep's DA part): Global heads (Deep flavour) + NonLocalAlignmentHead patch loss +
    InstanceAlignmentHead_DAF CE against the oracle heads / losses, fp32 engine."""
    from unsupervised_domain_adaptation_object_detection_implementation_b200 import hotpath
    uda.set_engine("simt_f32")
    m = hotpath.DeepHotPath((64, 64, 64), 1024).eval()
    seeded.fill_state_(m, 4, "deep.")
    c3, c4, c5 = (seeded.feature_map(f"deep.c{i}", (2, 64, h, w), 4) for i, (h, w) in enumerate(((8, 12), (12, 20), (13, 19))))
    feats = seeded.feature_map("deep.roi", (48, 1024), 4)
    gt = torch.tensor([0, 1])
    sd = {k: v.detach().double() for k, v in m.state_dict().items()}
    sub = lambda p: {k[len(p):]: v for k, v in sd.items() if k.startswith(p)}
    labels = (torch.arange(48) >= 24).long()
    ref = dict(
        globle_da_loss=0.1 * (da_oracle.ce2(da_oracle.global_alignment_head(c4.double(), sub("da_head_mid.")), gt) +
                              da_oracle.ce2(da_oracle.global_alignment_head(c5.double(), sub("da_head_top.")), gt)),
        patch_bottom_loss=0.1 * da_oracle.patch_loss(da_oracle.non_local_alignment_head(c3.double(), sub("local_da_head_bottom.")), gt),
        local_da_loss=0.2 * da_oracle.ce2(torch.sigmoid(da_oracle.instance_alignment_daf_logits(feats.double(), sub("local_da."))), labels))
    m = m.to(DEV)
    got = m.forward_train(c3.to(DEV), c4.to(DEV), c5.to(DEV), feats.to(DEV), [0, 1])
    for k in ref:
        assert abs(float(got[k]) - float(ref[k])) <= 2e-5 * abs(float(ref[k])), (k, float(got[k]), float(ref[k]))
