"""The detection stages on either side of the DA hot path, kept in plain PyTorch (adjacent subsystems,
SURVEY.md §2.1 / §8f rank 3) with the reference's class names, constructor arguments and return values so the
DA detectors build from `da_configs` and `forward_train` is functional:

  RPNHeadDA               mmdet/models/dense_heads/rpn_head_da.py:14-170   (loss on source images only, Q18)
  Shared2FCBBoxHead       mmdet/models/roi_heads/bbox_heads/convfc_bbox_head.py:198-253 (`forward_train_da`)
  StandardRoIHeadDA_v5    mmdet/models/roi_heads/standard_roi_head_da_v5.py:79-227

RoI pooling and the shared FCs run on libda_b200 (RoIAlign kernels, tensor-core GEMM); anchors, IoU assignment,
sampling and NMS (torchvision.ops.nms) are library/PyTorch code."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as F_
from .registry import HEADS, build_head, build_roi_extractor
from .roi_extractors import SingleRoIExtractor, bbox2roi  # noqa: F401  (SingleRoIExtractor registered below)

HEADS.register_module(module=SingleRoIExtractor)


# ------------------------------------------------------------------------------------------ core utilities
def bbox_overlaps(a, b):
    """IoU matrix [len(a), len(b)] of xyxy boxes."""
    if a.numel() == 0 or b.numel() == 0:
        return a.new_zeros(a.shape[0], b.shape[0])
    lt = torch.max(a[:, None, :2], b[None, :, :2])
    rb = torch.min(a[:, None, 2:4], b[None, :, 2:4])
    wh = (rb - lt).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    return inter / (area_a[:, None] + area_b[None, :] - inter).clamp(min=1e-6)


def max_iou_assign(boxes, gts, pos_iou_thr, neg_iou_thr, min_pos_iou, match_low_quality):
    """mmdet MaxIoUAssigner: returns assigned gt index per box (-1 ignore, 0 negative, k+1 positive for gt k)."""
    n = boxes.shape[0]
    assigned = boxes.new_full((n,), -1, dtype=torch.long)
    if gts.shape[0] == 0:
        assigned[:] = 0
        return assigned
    iou = bbox_overlaps(gts, boxes)
    max_iou, argmax = iou.max(dim=0)
    if isinstance(neg_iou_thr, (tuple, list)):
        assigned[(max_iou >= neg_iou_thr[0]) & (max_iou < neg_iou_thr[1])] = 0
    else:
        assigned[(max_iou >= 0) & (max_iou < neg_iou_thr)] = 0
    pos = max_iou >= pos_iou_thr
    assigned[pos] = argmax[pos] + 1
    if match_low_quality:
        gt_max, _ = iou.max(dim=1)
        for i in range(gts.shape[0]):
            if gt_max[i] >= min_pos_iou:
                assigned[iou[i] == gt_max[i]] = i + 1
    return assigned


def random_sample(assigned, num, pos_fraction, neg_pos_ub=-1):
    """mmdet RandomSampler: (pos_inds, neg_inds)."""
    pos = torch.nonzero(assigned > 0, as_tuple=False).squeeze(1)
    neg = torch.nonzero(assigned == 0, as_tuple=False).squeeze(1)
    n_pos = int(num * pos_fraction)
    if pos.numel() > n_pos:
        pos = pos[torch.randperm(pos.numel(), device=pos.device)[:n_pos]]
    n_neg = num - pos.numel()
    if neg_pos_ub >= 0:
        n_neg = min(n_neg, int(neg_pos_ub * max(1, pos.numel())))
    if neg.numel() > n_neg:
        neg = neg[torch.randperm(neg.numel(), device=neg.device)[:n_neg]]
    return pos, neg


class DeltaXYWHBBoxCoder:
    def __init__(self, target_means=(0., 0., 0., 0.), target_stds=(1., 1., 1., 1.), **kwargs):
        self.means, self.stds = target_means, target_stds

    def encode(self, boxes, gts):
        pw, ph = boxes[:, 2] - boxes[:, 0], boxes[:, 3] - boxes[:, 1]
        px, py = (boxes[:, 0] + boxes[:, 2]) * 0.5, (boxes[:, 1] + boxes[:, 3]) * 0.5
        gw, gh = gts[:, 2] - gts[:, 0], gts[:, 3] - gts[:, 1]
        gx, gy = (gts[:, 0] + gts[:, 2]) * 0.5, (gts[:, 1] + gts[:, 3]) * 0.5
        d = torch.stack([(gx - px) / pw, (gy - py) / ph, torch.log(gw / pw), torch.log(gh / ph)], -1)
        return (d - d.new_tensor(self.means)) / d.new_tensor(self.stds)

    def decode(self, boxes, deltas, max_shape=None, wh_ratio_clip=16 / 1000):
        d = deltas * deltas.new_tensor(self.stds) + deltas.new_tensor(self.means)
        max_ratio = abs(math.log(wh_ratio_clip))
        dw, dh = d[:, 2].clamp(-max_ratio, max_ratio), d[:, 3].clamp(-max_ratio, max_ratio)
        pw, ph = boxes[:, 2] - boxes[:, 0], boxes[:, 3] - boxes[:, 1]
        px, py = (boxes[:, 0] + boxes[:, 2]) * 0.5, (boxes[:, 1] + boxes[:, 3]) * 0.5
        gw, gh = pw * dw.exp(), ph * dh.exp()
        gx, gy = px + pw * d[:, 0], py + ph * d[:, 1]
        out = torch.stack([gx - gw * 0.5, gy - gh * 0.5, gx + gw * 0.5, gy + gh * 0.5], -1)
        if max_shape is not None:
            out[:, 0::2] = out[:, 0::2].clamp(0, max_shape[1])
            out[:, 1::2] = out[:, 1::2].clamp(0, max_shape[0])
        return out


class AnchorGenerator:
    """mmdet AnchorGenerator for one level (the DA configs use strides=[16])."""

    def __init__(self, strides, ratios, scales, **kwargs):
        self.strides = list(strides)
        ratios, scales = torch.tensor(ratios, dtype=torch.float32), torch.tensor(scales, dtype=torch.float32)
        self.base = []
        for s in self.strides:
            h_ratios = torch.sqrt(ratios)
            w_ratios = 1 / h_ratios
            ws = (s * w_ratios[:, None] * scales[None, :]).view(-1)
            hs = (s * h_ratios[:, None] * scales[None, :]).view(-1)
            self.base.append(torch.stack([-0.5 * ws, -0.5 * hs, 0.5 * ws, 0.5 * hs], -1))

    @property
    def num_base_anchors(self):
        return [b.shape[0] for b in self.base]

    def grid_anchors(self, featmap_size, level, device):
        h, w = featmap_size
        s = self.strides[level]
        sx = torch.arange(w, device=device, dtype=torch.float32) * s
        sy = torch.arange(h, device=device, dtype=torch.float32) * s
        yy, xx = torch.meshgrid(sy, sx, indexing="ij")
        shifts = torch.stack([xx, yy, xx, yy], -1).view(-1, 1, 4)
        return (shifts + self.base[level].to(device)[None]).view(-1, 4)


def smooth_l1(pred, target, beta=1.0):
    d = (pred - target).abs()
    return torch.where(d < beta, 0.5 * d * d / beta, d - 0.5 * beta)


# ------------------------------------------------------------------------------------------ RPN
@HEADS.register_module()
class RPNHeadDA(nn.Module):
    def __init__(self, in_channels, feat_channels=256, anchor_generator=None, bbox_coder=None, loss_cls=None,
                 loss_bbox=None, train_cfg=None, test_cfg=None, init_cfg=None, **kwargs):
        super().__init__()
        ag = dict(anchor_generator or dict(strides=[16], ratios=[0.5, 1.0, 2.0], scales=[8]))
        ag.pop("type", None)
        self.anchor_generator = AnchorGenerator(**ag)
        bc = dict(bbox_coder or {})
        bc.pop("type", None)
        self.bbox_coder = DeltaXYWHBBoxCoder(**bc)
        self.loss_cls_cfg, self.loss_bbox_cfg = dict(loss_cls or {}), dict(loss_bbox or {})
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        self.num_anchors = self.anchor_generator.num_base_anchors[0]
        self.rpn_conv = nn.Conv2d(in_channels, feat_channels, 3, padding=1)
        self.rpn_cls = nn.Conv2d(feat_channels, self.num_anchors, 1)
        self.rpn_reg = nn.Conv2d(feat_channels, self.num_anchors * 4, 1)
        for m in (self.rpn_conv, self.rpn_cls, self.rpn_reg):
            nn.init.normal_(m.weight, std=0.01)
            nn.init.constant_(m.bias, 0)

    def forward(self, feats):
        cls, reg = [], []
        for x in feats:
            x = F.relu(self.rpn_conv(x), inplace=True)
            cls.append(self.rpn_cls(x))
            reg.append(self.rpn_reg(x))
        return cls, reg

    def _proposals(self, cls, reg, anchors, img_shape, cfg):
        if cls.is_cuda:
            # native proposal stage (csrc/rpn_proposals.cu): scores, analytic anchors, decode, NMS, top max_per_img on the
            # device; the ONE host read is the survivor count, to hand back the reference's variable-length [n,5] tensor
            from . import functional as F_
            dets, count = F_.rpn_proposals(cls, reg, self.anchor_generator.base[0], self.anchor_generator.strides[0], img_shape,
                                           nms_pre=cfg.get("nms_pre", 2000), max_per_img=cfg.get("max_per_img", 1000),
                                           iou_thr=cfg.get("nms", {}).get("iou_threshold", 0.7),
                                           min_size=cfg.get("min_bbox_size", 0), means=self.bbox_coder.means, stds=self.bbox_coder.stds)
            return dets[:int(count)]
        scores = cls.permute(1, 2, 0).reshape(-1).sigmoid()
        deltas = reg.permute(1, 2, 0).reshape(-1, 4)
        nms_pre = cfg.get("nms_pre", 2000)
        if 0 < nms_pre < scores.numel():
            scores, idx = scores.topk(nms_pre)
            deltas, anchors = deltas[idx], anchors[idx]
        boxes = self.bbox_coder.decode(anchors, deltas, max_shape=img_shape)
        min_size = cfg.get("min_bbox_size", 0)
        if min_size >= 0:
            keep = ((boxes[:, 2] - boxes[:, 0]) > min_size) & ((boxes[:, 3] - boxes[:, 1]) > min_size)
            boxes, scores = boxes[keep], scores[keep]
        from torchvision.ops import nms
        keep = nms(boxes.float(), scores.float(), cfg.get("nms", {}).get("iou_threshold", 0.7))[:cfg.get("max_per_img", 1000)]
        return torch.cat([boxes[keep], scores[keep, None]], -1)

    def _loss_single_image(self, cls, reg, anchors, gt, img_shape):
        cfg = self.train_cfg
        a = cfg["assigner"]
        inside = (anchors[:, 0] >= -cfg.get("allowed_border", 0)) & (anchors[:, 1] >= -cfg.get("allowed_border", 0)) & \
                 (anchors[:, 2] < img_shape[1] + cfg.get("allowed_border", 0)) & (anchors[:, 3] < img_shape[0] + cfg.get("allowed_border", 0))
        idx_in = torch.nonzero(inside, as_tuple=False).squeeze(1)
        an = anchors[idx_in]
        assigned = max_iou_assign(an, gt, a["pos_iou_thr"], a["neg_iou_thr"], a.get("min_pos_iou", 0.0), a.get("match_low_quality", True))
        s = cfg["sampler"]
        pos, neg = random_sample(assigned, s["num"], s["pos_fraction"], s.get("neg_pos_ub", -1))
        scores = cls.permute(1, 2, 0).reshape(-1)[idx_in]
        deltas = reg.permute(1, 2, 0).reshape(-1, 4)[idx_in]
        sel = torch.cat([pos, neg])
        labels = torch.cat([scores.new_ones(pos.numel()), scores.new_zeros(neg.numel())])
        n_total = max(sel.numel(), 1)
        loss_cls = F.binary_cross_entropy_with_logits(scores[sel], labels, reduction="sum") / n_total
        if pos.numel():
            tgt = self.bbox_coder.encode(an[pos], gt[assigned[pos] - 1])
            loss_bbox = smooth_l1(deltas[pos], tgt, self.loss_bbox_cfg.get("beta", 1.0)).sum() / n_total
        else:
            loss_bbox = deltas.sum() * 0
        return loss_cls * self.loss_cls_cfg.get("loss_weight", 1.0), loss_bbox * self.loss_bbox_cfg.get("loss_weight", 1.0)

    def forward_train(self, x, img_metas, gt_bboxes, gt_da=None, gt_labels=None, gt_bboxes_ignore=None, proposal_cfg=None,
                      **kwargs):
        """-> (losses | None, proposal_list).  Loss on SOURCE images only, and only when the batch holds a target
        image (rpn_head_da.py:105-108,146-168, Q18)."""
        cls, reg = self(x)
        cls, reg = cls[0], reg[0]
        n, _, h, w = cls.shape
        anchors = self.anchor_generator.grid_anchors((h, w), 0, cls.device)
        gt_da = [int(d) for d in (gt_da.tolist() if torch.is_tensor(gt_da) else gt_da)] if gt_da is not None else [0] * n
        losses = None
        if any(d != 0 for d in gt_da):
            lc, lb, cnt = 0, 0, 0
            for i in range(n):
                if gt_da[i] == 0:
                    a, b = self._loss_single_image(cls[i], reg[i], anchors, gt_bboxes[i], img_metas[i]["img_shape"])
                    lc, lb, cnt = lc + a, lb + b, cnt + 1
            if cnt:
                losses = dict(loss_rpn_cls=lc / cnt, loss_rpn_bbox=lb / cnt)
        proposal_cfg = proposal_cfg if proposal_cfg is not None else (self.test_cfg or {})
        with torch.no_grad():
            props = [self._proposals(cls[i], reg[i], anchors, img_metas[i]["img_shape"], proposal_cfg) for i in range(n)]
        return losses, props


# ------------------------------------------------------------------------------------------ bbox head
@HEADS.register_module()
class Shared2FCBBoxHead(nn.Module):
    def __init__(self, in_channels=256, fc_out_channels=1024, roi_feat_size=7, num_classes=80, bbox_coder=None,
                 reg_class_agnostic=False, loss_cls=None, loss_bbox=None, num_shared_fcs=2, init_cfg=None, **kwargs):
        super().__init__()
        self.in_channels, self.fc_out_channels, self.num_classes = in_channels, fc_out_channels, num_classes
        self.roi_feat_size, self.reg_class_agnostic = roi_feat_size, reg_class_agnostic
        bc = dict(bbox_coder or dict(target_stds=(0.1, 0.1, 0.2, 0.2)))
        bc.pop("type", None)
        self.bbox_coder = DeltaXYWHBBoxCoder(**bc)
        self.loss_cls_cfg, self.loss_bbox_cfg = dict(loss_cls or {}), dict(loss_bbox or {})
        dims = [in_channels * roi_feat_size * roi_feat_size] + [fc_out_channels] * num_shared_fcs
        self.shared_fcs = nn.ModuleList([nn.Linear(dims[i], dims[i + 1]) for i in range(num_shared_fcs)])
        self.fc_cls = nn.Linear(fc_out_channels, num_classes + 1)
        self.fc_reg = nn.Linear(fc_out_channels, 4 if reg_class_agnostic else 4 * num_classes)
        for fc in self.shared_fcs:
            nn.init.xavier_uniform_(fc.weight)
            nn.init.constant_(fc.bias, 0)
        nn.init.normal_(self.fc_cls.weight, std=0.01)
        nn.init.normal_(self.fc_reg.weight, std=0.001)
        nn.init.constant_(self.fc_cls.bias, 0)
        nn.init.constant_(self.fc_reg.bias, 0)

    def forward_train_da(self, x):
        """-> (cls_score, bbox_pred, feat): `feat` [R, fc_out] feeds the instance-level domain classifier."""
        k = x.shape[0]
        t = F_.cast(x.flatten(1), F_.act_dtype()).view(k, 1, 1, -1)
        for fc in self.shared_fcs:
            t = F_.dense_layer(t, fc.weight, None, fc.bias, relu=True)
        feat = t.view(k, -1)
        f32 = feat.float()
        return self.fc_cls(f32), self.fc_reg(f32), feat

    def forward(self, x):
        cls, reg, _ = self.forward_train_da(x)
        return cls, reg

    def loss(self, cls_score, bbox_pred, labels, bbox_targets, pos_mask):
        """CrossEntropyLoss(use_sigmoid=True) over num_classes+1 channels + SmoothL1 on positives.  Both are element SUMS
        divided by the number of sampled RoIs: mmdet's binary_cross_entropy reduces with weight_reduce_loss(avg_factor =
        number of RoIs with label_weight > 0) (losses/cross_entropy_loss.py:100-114, bbox_heads/bbox_head.py:268-274), not
        by the R*(num_classes+1) elements a plain reduction='mean' would use."""
        n = max(cls_score.shape[0], 1)
        onehot = F.one_hot(labels, self.num_classes + 1).to(cls_score.dtype)
        out = dict(loss_cls=F.binary_cross_entropy_with_logits(cls_score, onehot, reduction="sum") / n *
                   self.loss_cls_cfg.get("loss_weight", 1.0),
                   acc=(cls_score.argmax(1) == labels).float().mean() * 100)
        if pos_mask.any():
            if self.reg_class_agnostic:
                pred = bbox_pred[pos_mask]
            else:
                pred = bbox_pred.view(bbox_pred.shape[0], -1, 4)[pos_mask, labels[pos_mask]]
            out["loss_bbox"] = smooth_l1(pred, bbox_targets[pos_mask], self.loss_bbox_cfg.get("beta", 1.0)).sum() / n * \
                self.loss_bbox_cfg.get("loss_weight", 1.0)
        else:
            out["loss_bbox"] = bbox_pred.sum() * 0
        return out


# ------------------------------------------------------------------------------------------ RoI head
@HEADS.register_module()
class StandardRoIHeadDA_v5(nn.Module):
    def __init__(self, bbox_roi_extractor=None, bbox_head=None, mask_roi_extractor=None, mask_head=None, shared_head=None,
                 train_cfg=None, test_cfg=None, pretrained=None, init_cfg=None):
        super().__init__()
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        self.bbox_roi_extractor = build_roi_extractor(bbox_roi_extractor)
        self.bbox_head = build_head(bbox_head)

    @property
    def with_bbox(self):
        return True

    def _sample(self, proposals, gt_bboxes, gt_labels):
        a, s = self.train_cfg["assigner"], self.train_cfg["sampler"]
        boxes = proposals[:, :4]
        if s.get("add_gt_as_proposals", True) and gt_bboxes.numel():
            boxes = torch.cat([gt_bboxes, boxes], 0)
        assigned = max_iou_assign(boxes, gt_bboxes, a["pos_iou_thr"], a["neg_iou_thr"], a.get("min_pos_iou", 0.0),
                                  a.get("match_low_quality", False))
        pos, neg = random_sample(assigned, s["num"], s["pos_fraction"], s.get("neg_pos_ub", -1))
        sel = torch.cat([pos, neg])
        labels = torch.full((sel.numel(),), self.bbox_head.num_classes, dtype=torch.long, device=boxes.device)
        targets = boxes.new_zeros(sel.numel(), 4)
        if pos.numel():
            labels[:pos.numel()] = gt_labels[assigned[pos] - 1]
            targets[:pos.numel()] = self.bbox_head.bbox_coder.encode(boxes[pos], gt_bboxes[assigned[pos] - 1])
        pos_mask = torch.zeros(sel.numel(), dtype=torch.bool, device=boxes.device)
        pos_mask[:pos.numel()] = True
        return boxes[sel], labels, targets, pos_mask

    def forward_train(self, x, img_metas, proposal_list, gt_bboxes, gt_labels, gt_da=None, gt_bboxes_ignore=None,
                      gt_masks=None, **kwargs):
        """-> (losses, bbox_feats=[feat_src, feat_tar], bbox_cls=[cls_src, cls_tar]); bbox loss on the source image
        only (standard_roi_head_da_v5.py:199-217).  RoIs of image i carry batch_ind = i and are pooled from the full
        [N,C,H,W] map (the intended semantics of the reference's sliced call, Q1)."""
        n = len(img_metas)
        sampled = [self._sample(proposal_list[i], gt_bboxes[i], gt_labels[i]) for i in range(n)]
        rois = bbox2roi([s[0] for s in sampled])
        roi_feats = self.bbox_roi_extractor(x[:self.bbox_roi_extractor.num_inputs], rois)
        cls_score, bbox_pred, feat = self.bbox_head.forward_train_da(roi_feats)
        sizes = [s[0].shape[0] for s in sampled]
        cls_l, reg_l, feat_l = cls_score.split(sizes), bbox_pred.split(sizes), feat.split(sizes)
        losses = dict()
        gt_da = [int(d) for d in (gt_da.tolist() if torch.is_tensor(gt_da) else gt_da)] if gt_da is not None else [0] * n
        src = [i for i in range(n) if gt_da[i] == 0]
        if src:
            i = src[0]
            losses.update(self.bbox_head.loss(cls_l[i], reg_l[i], sampled[i][1], sampled[i][2], sampled[i][3]))
        return losses, list(feat_l), list(cls_l)
