"""CPU suite (-m "not gpu"): pins the ORACLE against the golden vectors the reference's own
classes produced (tests/golden, generator oracle/make_golden.py), checks the closed-form losses
against literal transcriptions of the reference loops, and checks the C-ABI surface."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import da_oracle, roi_align as oracle_roi, seeded
from helpers import HEADS, build_head, rel_err, check_summary

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---------------------------------------------------------------- RoIAlign oracle
def test_roi_align_oracle_matches_torchvision_golden(golden):
    g = golden("roi_align_torchvision.pt")
    out, grid, bidx = oracle_roi.roi_align_forward(g["feat"].numpy(), g["rois"].numpy(), 7, 1.0 / g["stride"], 0, True)
    ref = g["out"].numpy()
    assert np.abs(out - ref).max() <= 1e-6 * np.abs(ref).max()
    assert np.array_equal(grid, g["grid"].numpy())            # sampling grid: bit-exact
    assert np.array_equal(bidx, g["batch_idx"].numpy())       # batch indices: bit-exact
    gin = oracle_roi.roi_align_backward(g["cot"].numpy(), g["rois"].numpy(), tuple(g["feat"].shape), 7, 1.0 / g["stride"], 0, True)
    dref = g["dfeat"].numpy()
    assert np.abs(gin - dref).max() <= 1e-5 * np.abs(dref).max()


def test_roi_align_oracle_legacy_mode(golden):
    g = golden("roi_align_torchvision.pt")
    out, _, _ = oracle_roi.roi_align_forward(g["feat"].numpy(), g["rois"][:24].numpy(), 7, 1.0 / g["stride"], 2, False)
    ref = g["out_legacy_sr2"].numpy()
    assert np.abs(out - ref).max() <= 1e-6 * np.abs(ref).max()


def test_roi_align_oracle_live_torchvision():
    torchvision = pytest.importorskip("torchvision")
    feat = seeded.seeded_tensor("live.feat", (3, 5, 17, 23), 1)
    rois = torch.cat([seeded.synthetic_rois(40, 3, 17 * 8, 23 * 8, 1, 4.0, 200.0), seeded.adversarial_rois(3, 17 * 8, 23 * 8)])
    ref = torch.ops.torchvision.roi_align(feat, rois, 1.0 / 8, 7, 7, 0, True).numpy()
    out, _, _ = oracle_roi.roi_align_forward(feat.numpy(), rois.numpy(), 7, 1.0 / 8, 0, True)
    assert np.abs(out - ref).max() <= 1e-6 * np.abs(ref).max()


def test_roi_align_oracle_edge_cases():
    feat = np.ones((1, 2, 8, 8), np.float32)
    # empty RoI set, batch index out of range (Q1: bounds-checked -> zeros), zero-area RoI
    out, grid, _ = oracle_roi.roi_align_forward(feat, np.zeros((0, 5), np.float32), 7, 0.25)
    assert out.shape == (0, 2, 7, 7)
    rois = np.array([[3, 0, 0, 16, 16], [0, 8, 8, 8, 8], [0, 4, 4, 20, 20]], np.float32)
    out, grid, bidx = oracle_roi.roi_align_forward(feat, rois, 7, 0.25)
    assert np.all(out[0] == 0) and np.all(out[1] == 0)
    assert grid[1].tolist() == [0, 0]
    assert np.allclose(out[2], 1.0)


def test_map_roi_levels_matches_reference_formula():
    rois = seeded.synthetic_rois(200, 1, 1024, 2048, 3, 8.0, 900.0)
    scale = torch.sqrt((rois[:, 3] - rois[:, 1]) * (rois[:, 4] - rois[:, 2]))
    ref = torch.floor(torch.log2(scale / 56 + 1e-6)).clamp(min=0, max=3).long()  # single_level_roi_extractor.py:51-54
    got = oracle_roi.map_roi_levels(rois.numpy(), 4, 56.0)
    assert np.array_equal(got, ref.numpy())


# ---------------------------------------------------------------- heads oracle vs reference golden
@pytest.mark.parametrize("name", sorted(HEADS))
def test_head_oracle_matches_reference_golden(golden, name):
    g = golden(f"head_{name}.pt")
    m = build_head(name, g["seed"])          # THIS repo's module: same state_dict keys as the reference class
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in m.state_dict().items()}
    x = g["x"].clone().requires_grad_(True)
    out = HEADS[name][1](x, sd)
    assert rel_err(out, g["out"][0]) <= 2e-6
    (out * g["cot"][0]).sum().backward()
    assert rel_err(x.grad, g["dx"]) <= 2e-5
    for k, summ in g["dparams"].items():
        assert sd[k].grad is not None, k
        assert check_summary(sd[k].grad, summ, 1e-4) <= 1e-4, k
    # parameters the reference never reaches (dead branch Q10, unused BN Q9) get no gradient there either
    for k in g["no_grad_params"]:
        assert k in sd


def test_state_dict_keys_match_reference_surface():
    """parameter names/shapes of the DA modules."""
    m = build_head("global_alignment_cbam")
    sd = m.state_dict()
    assert sd["conv1.weight"].shape == (32, 64, 3, 3) and sd["CBAM.conv.weight"].shape == (1, 2, 7, 7)
    assert sd["CBAM.mlp.0.weight"].shape == (2, 32, 1, 1) and sd["fc2.weight"].shape == (2, 8)
    m = build_head("instance_alignment")
    sd = m.state_dict()
    assert sd["nlb.conv_phi.weight"].shape == (512, 1024, 1, 1) and sd["fc3.weight"].shape == (2, 512)
    assert "bn1.running_mean" in sd and "bn2.weight" in sd
    assert set(p for p, _ in m.named_parameters() if p.startswith("bn")) == {"bn1.weight", "bn1.bias", "bn2.weight", "bn2.bias"}
    m = build_head("srm")
    assert m.state_dict()["conv2.weight"].shape == (144, 16, 3, 3) and m.conv1.padding == (1, 1) and m.conv2.padding == (3, 3)


# ---------------------------------------------------------------- loss tails vs the reference backbones
def test_loss_tails_match_reference_backbones(golden):
    g = golden("backbone_loss_tails.pt")
    d = g["daf_org"]
    assert abs(float(da_oracle.daf_image_loss(d["patch_feat"], d["gt"])) - float(d["loss"])) <= 1e-6
    assert abs(float(da_oracle.daf_image_loss_loop(d["patch_feat"], d["gt"])) - float(d["loss"])) <= 1e-6
    assert abs(float(da_oracle.daf_image_loss(d["patch_feat"], g["daf_org_tt"]["gt"])) - float(g["daf_org_tt"]["loss"])) <= 1e-6
    m = g["maf"]
    for pred, ref in zip(m["preds"], m["losses"]):
        assert abs(float(da_oracle.ce2(pred, m["gt"])) - float(ref)) <= 1e-6   # CE on sigmoid outputs (Q4)
    c = g["cbam"]
    for z, ref in zip(c["logits"], c["global_losses"]):
        assert abs(float(da_oracle.ce2(z, c["gt"])) - float(ref)) <= 1e-6
    assert abs(float(da_oracle.patch_loss(c["local_feat"], c["gt"])) - float(c["patch_loss"])) <= 1e-6
    assert abs(float(da_oracle.patch_loss_loop(c["local_feat"], c["gt"])) - float(c["patch_loss"])) <= 1e-6


def test_focal_matches_reference(golden):
    g = golden("focal_loss.pt")
    u = g["u"].clone().requires_grad_(True)
    loss = da_oracle.focal2(u, g["labels"])
    assert abs(float(loss) - float(g["loss"])) <= 1e-6
    loss.backward()
    assert rel_err(u.grad, g["du"]) <= 1e-5


def test_reference_known_answer_ce():
    """The only golden numbers the reference's tests hold near the path
    (tests/test_metrics/test_losses.py:18-33): CE([[100,-100]], label 1) == 200."""
    assert abs(float(da_oracle.ce2(torch.tensor([[100.0, -100.0]]), torch.tensor([1]))) - 200.0) < 1e-4


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_closed_forms_match_loop_transcriptions(seed):
    img = seeded.feature_map("cf.img", (2, 1, 5, 7), seed).double()
    pred = torch.sigmoid(seeded.seeded_tensor("cf.pred", (30, 2), seed).double())
    labels = torch.cat([torch.zeros(13), torch.ones(17)]).long()
    a = da_oracle.consistency_loss(img, pred, labels)
    b = da_oracle.consistency_loss_loop(img, pred, labels)
    assert abs(float(a - b)) <= 1e-10
    for gt in ([0, 1], [1, 0], [0, 0], [1, 1]):
        gt = torch.tensor(gt)
        p = seeded.seeded_tensor("cf.p", (2, 1, 5, 7), seed).double()
        assert abs(float(da_oracle.daf_image_loss(p, gt) - da_oracle.daf_image_loss_loop(p, gt))) <= 1e-12
        assert abs(float(da_oracle.patch_loss(p, gt) - da_oracle.patch_loss_loop(p, gt))) <= 1e-12


def test_bbox2roi_batch_indices():
    boxes = [torch.rand(5, 4), torch.zeros(0, 4), torch.rand(3, 5)]
    from unsupervised_domain_adaptation_object_detection_implementation_b200.roi_extractors import bbox2roi, bbox2roi_train
    rois = bbox2roi(boxes)
    assert torch.equal(rois, da_oracle.bbox2roi(boxes))
    assert rois[:, 0].tolist() == [0.0] * 5 + [2.0] * 3
    per = bbox2roi_train(boxes)
    assert [len(p) for p in per] == [5, 0, 3] and torch.equal(torch.cat(per), rois)


# ---------------------------------------------------------------- C-ABI surface
def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "da_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(da_[a-z0-9_]+)\s*\(", hdr)))


def test_c_abi_exports_every_declared_symbol():
    from unsupervised_domain_adaptation_object_detection_implementation_b200 import _lib
    syms = _declared_symbols()
    assert len(syms) >= 30
    so = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in syms if not hasattr(so, s)]
    assert not missing, missing
    assert sorted(_lib.SIGNATURES) == syms          # the ctypes table covers the header exactly
    assert so.da_version() >= 100


def test_product_path_refuses_cpu_tensors():
    from unsupervised_domain_adaptation_object_detection_implementation_b200 import functional as F_
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        F_.roi_align(torch.zeros(1, 4, 8, 8), torch.zeros(1, 5), 7, 1.0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        F_.pixel_domain_loss(torch.zeros(2, 1, 4, 4), torch.tensor([0, 1]), True)


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "unsupervised_domain_adaptation_object_detection_implementation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f


def test_group_local_da_loss_oracle_matches_reference_methods(golden):
    """L5: the restatement against the values the reference's own methods returned (oracle/make_golden.group_loss_cases)."""
    from helpers import group_case, group_heads
    g = golden("group_local_da_loss.pt")
    assert {r["flavour"] for r in g.values()} == {"daf", "maf", "deep"}
    for name, rec in g.items():
        feats, cls = group_case(rec, name)
        fore, back = group_heads(rec, name)
        torch.manual_seed(rec["rng_seed"])        # the centroid draws of the DAF flavour come from the global RNG
        val = da_oracle.group_local_da_loss(feats, cls, fore.state_dict(), back.state_dict(), rec["flavour"])
        assert abs(float(val) - rec["loss"]) <= 2e-6 * max(1.0, abs(rec["loss"])), (name, float(val), rec["loss"])


def test_bbox_head_loss_matches_reference_loss_classes(golden):
    """R3's box-head loss: the oracle AND the product's Shared2FCBBoxHead.loss (plain torch, device-agnostic) against the
    values and gradients of the reference's own CrossEntropyLoss(use_sigmoid=True) / SmoothL1Loss / accuracy
    (oracle/make_golden.bbox_head_loss_case).  The classification loss is a SUM over the (num_classes+1) channels divided by
    the number of RoIs -- not a mean over elements."""
    from unsupervised_domain_adaptation_object_detection_implementation_b200 import detection
    g = golden("bbox_head_loss.pt")
    C = g["num_classes"]
    head = detection.Shared2FCBBoxHead(in_channels=4, fc_out_channels=8, num_classes=C)
    for fn in (lambda a, b: da_oracle.bbox_head_loss(a, b, g["labels"], g["targets"], g["pos_mask"], C),
               lambda a, b: head.loss(a, b, g["labels"], g["targets"], g["pos_mask"])):
        cls = g["cls_score"].clone().requires_grad_(True)
        reg = g["bbox_pred"].clone().requires_grad_(True)
        out = fn(cls, reg)
        assert abs(float(out["loss_cls"]) - float(g["loss_cls"])) <= 1e-6 * float(g["loss_cls"])
        assert abs(float(out["loss_bbox"]) - float(g["loss_bbox"])) <= 1e-6 * float(g["loss_bbox"])
        assert abs(float(out["acc"]) - float(g["acc"])) <= 1e-4
        (out["loss_cls"] + out["loss_bbox"]).backward()
        assert float((cls.grad - g["dcls"]).abs().max()) <= 1e-7 and float((reg.grad - g["dreg"]).abs().max()) <= 1e-7


# ---------------------------------------------------------------------------- RPN proposal stage
def _rpn_inputs(name, g):
    from oracle import seeded
    cls = seeded.seeded_tensor(f"rpn.{name}.cls", (g["A"], g["H"], g["W"]), 0, scale=g["cls_scale"])
    reg = seeded.seeded_tensor(f"rpn.{name}.reg", (4 * g["A"], g["H"], g["W"]), 0, scale=g["reg_scale"])
    return cls, reg


def test_rpn_oracle_vs_reference_golden(golden):
    """oracle/rpn_oracle.py against the fixture produced by the reference's AnchorGenerator + delta2bbox run in place
    (oracle/make_golden_rpn.py) and torchvision's NMS: bit-exact anchors, ranking, decoded boxes and proposals."""
    from oracle import rpn_oracle
    rec = golden("rpn_proposals.pt")
    d = rec["docstring"]
    out = rpn_oracle.delta2bbox(d["rois"], d["deltas"], max_shape=d["max_shape"])
    assert torch.equal(out, d["reference_output"])
    assert torch.allclose(out, d["expected"], atol=5e-5)          # the known-answer example of delta_xywh_bbox_coder.py:210-222
    for name in ("small", "min_size", "dc5_s"):
        g = rec[name]
        cls, reg = _rpn_inputs(name, g)
        base = rpn_oracle.base_anchors(g["stride"], [0.5, 1.0, 2.0], [2, 4, 8, 16, 32])
        assert torch.equal(base, g["base_anchors"])
        anchors = rpn_oracle.grid_anchors(base, g["H"], g["W"], g["stride"])
        assert torch.equal(anchors[:32], g["anchors_first"]) and torch.equal(anchors[-32:], g["anchors_last"])
        s, idx = rpn_oracle.rank(cls.permute(1, 2, 0).reshape(-1).sigmoid(), g["nms_pre"])
        assert torch.equal(idx[:64], g["top_idx_head"]) and idx.numel() == g["n_decoded"]
        decoded = rpn_oracle.delta2bbox(anchors[idx], reg.permute(1, 2, 0).reshape(-1, 4)[idx], max_shape=g["img_shape"])
        assert torch.equal(decoded[:256], g["decoded_head"])
        dets = rpn_oracle.proposals(cls, reg, base, g["stride"], g["img_shape"], g["nms_pre"], g["max_per_img"], g["iou_thr"], g["min_size"])
        assert dets.shape == g["dets"].shape and torch.equal(dets, g["dets"]), name


def test_rpn_oracle_nms_vs_torchvision():
    """The greedy NMS loop against torchvision.ops.nms (the stand-in for mmcv.ops.nms) on clustered random boxes."""
    from torchvision.ops import nms
    from oracle import rpn_oracle
    g = torch.Generator().manual_seed(3)
    ctr = torch.rand(40, 2, generator=g) * 400
    xy = (ctr[torch.randint(0, 40, (3000,), generator=g)] + torch.randn(3000, 2, generator=g) * 12)
    wh = torch.rand(3000, 2, generator=g) * 80 + 4
    boxes = torch.cat([xy - wh / 2, xy + wh / 2], -1)
    scores = torch.sort(torch.rand(3000, generator=g), descending=True).values
    for thr in (0.3, 0.7):
        assert torch.equal(rpn_oracle.greedy_nms(boxes, thr), nms(boxes, scores, thr))
