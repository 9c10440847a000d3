"""GPU parity of the RPN proposal stage (csrc/rpn_proposals.cu, functional.rpn_proposals) against
oracle/rpn_oracle.py and the fixture tests/golden/rpn_proposals.pt (the reference's own AnchorGenerator + delta2bbox run in place,
torchvision NMS for the absent mmcv op).

Bars: ranking order and NMS keep set are INDEX work: bit-exact against the oracle given the same scores / boxes; decoded boxes are
fp32 with the reference's rounding points: <= 1e-6 of the image size against the CPU oracle (expf differs between the CPU's and
the GPU's math library in the last ulp), bit-exact against the same arithmetic done by torch ON the device."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from unsupervised_domain_adaptation_object_detection_implementation_b200 import functional as F_, detection  # noqa: E402
from oracle import rpn_oracle, seeded  # noqa: E402

DEV = "cuda"
RATIOS, SCALES = [0.5, 1.0, 2.0], [2, 4, 8, 16, 32]


def _inputs(name, g):
    cls = seeded.seeded_tensor(f"rpn.{name}.cls", (g["A"], g["H"], g["W"]), 0, scale=g["cls_scale"])
    reg = seeded.seeded_tensor(f"rpn.{name}.reg", (4 * g["A"], g["H"], g["W"]), 0, scale=g["reg_scale"])
    return cls, reg


@pytest.mark.parametrize("name", ["small", "min_size", "dc5_s"])
def test_rpn_proposals_vs_reference_golden(golden, name):
    g = golden("rpn_proposals.pt")[name]
    cls, reg = _inputs(name, g)
    base = rpn_oracle.base_anchors(g["stride"], RATIOS, SCALES)
    dets, count, keep = F_.rpn_proposals(cls.to(DEV), reg.to(DEV), base, g["stride"], g["img_shape"], g["nms_pre"], g["max_per_img"],
                                         g["iou_thr"], g["min_size"], return_keep=True)
    n = int(count)
    # (1) the scores and their ranking: sigmoid on the device vs the CPU's (same formula; the order must agree wherever the
    # CPU scores are not tied)
    dev_scores = torch.empty(g["A"] * g["H"] * g["W"], device=DEV)
    F_.check(F_.lib.da_rpn_scores(F_._ptr(cls.to(DEV).contiguous()), g["A"], g["H"] * g["W"], F_._ptr(dev_scores), F_._stream()), "scores")
    cpu_scores = cls.permute(1, 2, 0).reshape(-1).sigmoid()
    assert float((dev_scores.cpu() - cpu_scores).abs().max()) <= 2.4e-7
    assert torch.equal(dev_scores, torch.sigmoid(cls.to(DEV).permute(1, 2, 0).reshape(-1)))            # bit-exact vs torch on the device
    # (2) decoded boxes (pre-NMS set, rank order) vs the oracle fed with the DEVICE's scores (identical ranking)
    n_pre = g["n_decoded"]
    boxes_dev, valid_dev = F_.rpn_decoded_boxes(n_pre, torch.device(DEV))
    s, idx = rpn_oracle.rank(dev_scores.cpu(), g["nms_pre"])
    anchors = rpn_oracle.grid_anchors(base, g["H"], g["W"], g["stride"])
    ref_boxes = rpn_oracle.delta2bbox(anchors[idx], reg.permute(1, 2, 0).reshape(-1, 4)[idx], max_shape=g["img_shape"])
    assert float((boxes_dev.cpu() - ref_boxes).abs().max()) <= 1e-6 * max(g["img_shape"]) * 4
    # (3) index work: validity + greedy NMS on the DEVICE's boxes, bit-exact keep set and order
    b = boxes_dev.cpu()
    valid = ((b[:, 2] - b[:, 0]) > g["min_size"]) & ((b[:, 3] - b[:, 1]) > g["min_size"]) if g["min_size"] >= 0 else torch.ones(n_pre, dtype=torch.bool)
    assert torch.equal(valid, valid_dev.cpu())
    vi = torch.nonzero(valid).squeeze(1)
    ref_keep = vi[rpn_oracle.greedy_nms(b[vi], g["iou_thr"])][:g["max_per_img"]]
    assert n == ref_keep.numel()
    assert torch.equal(keep[:n].cpu().long(), ref_keep)
    assert bool((keep[n:] == -1).all()) and bool((dets[n:] == 0).all())
    assert torch.equal(dets[:n, :4].cpu(), b[ref_keep]) and torch.equal(dets[:n, 4].cpu(), s[ref_keep])
    # (4) end to end against the reference-made fixture
    # (the CPU's and the device's sigmoid / exp may differ in the last ulp, which can swap two neighbours in the ranking or flip
    # an IoU that sits on the threshold: compare as sets, allow 1 % of such flips; steps (1)-(3) pin the arithmetic itself)
    gd = g["dets"]
    assert abs(n - gd.shape[0]) <= max(2, gd.shape[0] // 100)
    dist = torch.cdist(dets[:n, :4].cpu().double(), gd[:, :4].double(), p=float("inf"))
    tol = 1e-6 * max(g["img_shape"]) * 4
    matched = (dist.min(1).values <= tol).float().mean().item(), (dist.min(0).values <= tol).float().mean().item()
    assert min(matched) >= 0.99, matched


def test_rpn_proposals_full_size_and_edge_cases():
    """C5 of a 1024x2048 input (64x128 cells x 15 anchors = 122 880 candidates, nms_pre 12000, max_per_img 2000: the train
    configuration of faster_rcnn_r50_torch_daf.py:78-82): index work bit-exact against the oracle on the device's boxes; plus
    nms_pre >= candidates, everything filtered out, and a zero-candidate call."""
    A, H, W = 15, 64, 128
    cls = seeded.seeded_tensor("rpn.full.cls", (A, H, W), 0, scale=2.0).to(DEV)
    reg = seeded.seeded_tensor("rpn.full.reg", (4 * A, H, W), 0, scale=0.5).to(DEV)
    base = rpn_oracle.base_anchors(16, RATIOS, SCALES)
    dets, count, keep = F_.rpn_proposals(cls, reg, base, 16, (1024, 2048), 12000, 2000, 0.7, 0.0, return_keep=True)
    n = int(count)
    boxes, valid = F_.rpn_decoded_boxes(12000, torch.device(DEV))
    b = boxes.cpu()
    vi = torch.nonzero(valid.cpu()).squeeze(1)
    ref_keep = vi[rpn_oracle.greedy_nms(b[vi], 0.7)][:2000]
    assert n == ref_keep.numel() and torch.equal(keep[:n].cpu().long(), ref_keep)
    assert bool((dets[:n - 1, 4] >= dets[1:n, 4]).all())                                    # rank order
    assert bool((dets[:n, 0] >= 0).all()) and bool((dets[:n, 2] <= 2048).all()) and bool((dets[:n, 3] <= 1024).all())
    # torchvision's CUDA NMS on the same boxes agrees as well
    from torchvision.ops import nms
    s = torch.sort(torch.sigmoid(cls.permute(1, 2, 0).reshape(-1)), descending=True, stable=True).values[:12000]
    tv = nms(boxes[valid], s[valid], 0.7)[:2000]
    assert torch.equal(vi.to(DEV)[tv].cpu(), ref_keep)
    # nms_pre larger than the candidate count: all candidates ranked
    small_c, small_r = cls[:, :4, :5].contiguous(), reg[:, :4, :5].contiguous()
    d2, c2 = F_.rpn_proposals(small_c, small_r, base, 16, (64, 80), 12000, 50, 0.7, 0.0)
    ref = rpn_oracle.proposals(small_c.cpu(), small_r.cpu(), base, 16, (64, 80), 12000, 50, 0.7, 0.0,
                               scores=torch.sigmoid(small_c.permute(1, 2, 0).reshape(-1)).cpu())
    assert int(c2) == ref.shape[0] and float((d2[:int(c2)].cpu() - ref).abs().max()) <= 1e-3
    # min size nothing can pass: zero proposals, like proposals.new_zeros(0, 5)
    d3, c3 = F_.rpn_proposals(small_c, small_r, base, 16, (64, 80), 100, 50, 0.7, 1e6)
    assert int(c3) == 0 and bool((d3 == 0).all())
    with pytest.raises(RuntimeError):
        F_.rpn_proposals(small_c, reg, base, 16, (64, 80))


def test_rpn_head_uses_the_native_stage_and_matches_its_torch_path():
    """RPNHeadDA._proposals on CUDA tensors = the native stage; its CPU branch (torch ops + torchvision NMS, the r01 path) gives the
    same proposals on the same maps."""
    torch.manual_seed(0)
    head = detection.RPNHeadDA(64, feat_channels=64, anchor_generator=dict(type="AnchorGenerator", scales=SCALES, ratios=RATIOS, strides=[16]),
                               bbox_coder=dict(type="DeltaXYWHBBoxCoder", target_means=[0., 0., 0., 0.], target_stds=[1., 1., 1., 1.]))
    for m in (head.rpn_cls, head.rpn_reg):
        torch.nn.init.normal_(m.weight, std=0.3)
    x = seeded.seeded_tensor("rpn.head.x", (1, 64, 20, 28), 0)
    cfg = dict(nms_pre=1500, max_per_img=200, nms=dict(type="nms", iou_threshold=0.7), min_bbox_size=0)
    with torch.no_grad():
        cls, reg = head([x])
        anchors = head.anchor_generator.grid_anchors((20, 28), 0, "cpu")
        ref = head._proposals(cls[0][0], reg[0][0], anchors, (320, 448), cfg)
        got = head._proposals(cls[0][0].to(DEV), reg[0][0].to(DEV), anchors.to(DEV), (320, 448), cfg)
    assert got.is_cuda and got.shape == ref.shape
    assert float((got.cpu() - ref).abs().max()) <= 2e-3
