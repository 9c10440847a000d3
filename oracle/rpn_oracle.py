"""ORACLE — test infrastructure only (imported by tests/ and oracle/make_golden_rpn.py, never by the product).

CPU restatement of the RPN proposal stage of one image / one level:
  RPNHeadDA._get_bboxes_single + _bbox_post_process      mmdet/models/dense_heads/rpn_head_da.py:170-303
  AnchorGenerator (base anchors, grid)                    mmdet/core/anchor/anchor_generator.py:131-194,347-390
  delta2bbox                                              mmdet/core/bbox/coder/delta_xywh_bbox_coder.py:224-259
  batched_nms on one level = greedy NMS                   mmcv-full 1.3.17 (absent; pinned to torchvision.ops.nms, same lineage)
Pinned by tests/test_oracle_cpu.py against tests/golden/rpn_proposals.pt, which oracle/make_golden_rpn.py produced by running the
reference's own AnchorGenerator and delta2bbox in place + torchvision.ops.nms, and against delta2bbox's docstring example.
"""
import math

import numpy as np
import torch


def base_anchors(stride, ratios, scales):
    """gen_single_level_base_anchors (anchor_generator.py:151-194), scale_major=True, center = 0 (center_offset 0)."""
    ratios, scales = torch.tensor(ratios, dtype=torch.float32), torch.tensor(scales, dtype=torch.float32)
    h_ratios = torch.sqrt(ratios)
    w_ratios = 1 / h_ratios
    ws = (stride * w_ratios[:, None] * scales[None, :]).view(-1)
    hs = (stride * h_ratios[:, None] * scales[None, :]).view(-1)
    return torch.stack([-0.5 * ws, -0.5 * hs, 0.5 * ws, 0.5 * hs], -1)


def grid_anchors(base, H, W, stride):
    """single_level_grid_anchors (anchor_generator.py:347-390): index = (y * W + x) * A + a."""
    sx = torch.arange(W, dtype=torch.float32) * stride
    sy = torch.arange(H, dtype=torch.float32) * stride
    yy, xx = torch.meshgrid(sy, sx, indexing="ij")
    shifts = torch.stack([xx, yy, xx, yy], -1).view(-1, 1, 4)
    return (shifts + base[None]).view(-1, 4)


def delta2bbox(rois, deltas, means=(0., 0., 0., 0.), stds=(1., 1., 1., 1.), max_shape=None, wh_ratio_clip=16 / 1000):
    """delta_xywh_bbox_coder.py:224-259 (clip_border=True, add_ctr_clamp=False), fp32, same operation order."""
    if deltas.shape[0] == 0:
        return deltas
    d = deltas * deltas.new_tensor(stds).view(1, -1) + deltas.new_tensor(means).view(1, -1)
    dxy, dwh = d[:, :2], d[:, 2:]
    pxy = (rois[:, :2] + rois[:, 2:]) * 0.5
    pwh = rois[:, 2:] - rois[:, :2]
    dxy_wh = pwh * dxy
    max_ratio = np.abs(np.log(wh_ratio_clip))
    dwh = dwh.clamp(min=-max_ratio, max=max_ratio)
    gxy = pxy + dxy_wh
    gwh = pwh * dwh.exp()
    out = torch.cat([gxy - gwh * 0.5, gxy + gwh * 0.5], dim=-1)
    if max_shape is not None:
        out[:, 0::2].clamp_(min=0, max=max_shape[1])
        out[:, 1::2].clamp_(min=0, max=max_shape[0])
    return out


def greedy_nms(boxes, thr):
    """Boxes in rank order (best first) -> indices kept.  IoU = inter / (Sa + Sb - inter) in fp32, suppressed when > thr
    (mmcv nms_cuda / torchvision nms kernel: offset 0)."""
    b = boxes.detach().cpu().numpy().astype(np.float32)
    n = b.shape[0]
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    removed = np.zeros(n, dtype=bool)
    keep = []
    for i in range(n):
        if removed[i]:
            continue
        keep.append(i)
        if i + 1 == n:
            break
        r = b[i + 1:]
        w = np.maximum(np.minimum(b[i, 2], r[:, 2]) - np.maximum(b[i, 0], r[:, 0]), np.float32(0))
        h = np.maximum(np.minimum(b[i, 3], r[:, 3]) - np.maximum(b[i, 1], r[:, 1]), np.float32(0))
        inter = (w * h).astype(np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            iou = inter / ((area[i] + area[i + 1:]).astype(np.float32) - inter)
        removed[i + 1:] |= iou > np.float32(thr)
    return torch.tensor(keep, dtype=torch.long)


def rank(scores, nms_pre):
    """scores.sort(descending=True)[:nms_pre] (rpn_head_da.py:236-243); ties broken by index (stable), which the reference leaves open."""
    s, idx = torch.sort(scores, descending=True, stable=True)
    n = nms_pre if 0 < nms_pre < scores.numel() else scores.numel()
    return s[:n], idx[:n]


def proposals(cls, reg, base, stride, img_shape, nms_pre, max_per_img, iou_thr=0.7, min_size=0.0, means=(0., 0., 0., 0.),
              stds=(1., 1., 1., 1.), scores=None):
    """cls [A,H,W], reg [4A,H,W] -> dets [n,5].  `scores` (optional, [H*W*A]) overrides sigmoid(cls) (e.g. the device's values,
    to compare index work bit for bit)."""
    A, H, W = cls.shape
    if scores is None:
        scores = cls.permute(1, 2, 0).reshape(-1).sigmoid()
    deltas = reg.permute(1, 2, 0).reshape(-1, 4)
    anchors = grid_anchors(base, H, W, stride)
    s, idx = rank(scores, nms_pre)
    boxes = delta2bbox(anchors[idx], deltas[idx], means, stds, max_shape=img_shape)
    if min_size >= 0:
        valid = ((boxes[:, 2] - boxes[:, 0]) > min_size) & ((boxes[:, 3] - boxes[:, 1]) > min_size)
        boxes, s = boxes[valid], s[valid]
    if boxes.numel() == 0:
        return boxes.new_zeros(0, 5)
    keep = greedy_nms(boxes, iou_thr)[:max_per_img]
    return torch.cat([boxes[keep], s[keep, None]], -1)
