"""Domain losses of the DA path as the reference composes them.

  daf_image_loss          L1  mmdet/models/backbones/resnet_da_daf_org.py:816-822
  patch_loss              L2  mmdet/models/backbones/resnet_da_cbam.py:971-979
  image_ce_loss           L3  resnet_da_cbam.py:966-968 (raw logits), resnet_da.py:846-848 (on sigmoid)
  instance_ce_loss        L4  mmdet/models/detectors/DAFaster_rcnn_Orig.py:177-188
  group_local_da_loss     L5  DAFaster_rcnn.py:198-327, MAFaster_rcnn.py:204-299, DAFaster_rcnn_Deep.py:232-329
  FocalLoss               L6  mmdet/models/losses/focal_loss.py:106-182
  consistency_loss        L7  DAFaster_rcnn_Orig.py:161-175
"""
import torch
import torch.nn as nn

from . import functional as F_


def daf_image_loss(patch_feat, gt_domain):
    """L1 (Q5): every batch slot averages over the WHOLE [N,1,H,W] map."""
    return F_.pixel_domain_loss(patch_feat, gt_domain, whole_batch=True)


def patch_loss(local_feat, gt_domain):
    """L2: per-image mean, summed over images."""
    return F_.pixel_domain_loss(local_feat, gt_domain, whole_batch=False)


def image_ce_loss(z, gt_domain, on_sigmoid):
    """L3: returns (loss, pred) with pred = sigmoid(z) when on_sigmoid (SRM) else unused."""
    return F_.ce2(z, gt_domain, on_sigmoid)


def instance_ce_loss(z, labels):
    """L4: CE applied to sigmoid(fc3) (Q4); returns (loss, pred_da = sigmoid(z))."""
    return F_.ce2(z, labels, True)


def consistency_loss(imgs_feat, ins_preds, ins_labels):
    """L7.  ins_preds are the sigmoid outputs of the instance head (sigmoid is applied again
    inside, as in the reference)."""
    return F_.consistency_loss(imgs_feat, ins_preds, ins_labels)


def _draw_centroids(dim, device):
    """cluster.py:95-98: ten separate randn([dim]) draws (the reference draws them on the GPU)."""
    return torch.stack([torch.randn([dim], device=device) for _ in range(10)], 0)


@torch.no_grad()
def group_local_da_loss(bbox_feats, bbox_cls, head_fore, head_back, flavour="daf", k=20, draw_centroids=None):
    """L5, value-faithful to what the reference code computes (each point is spelled out with file:line in
    oracle/da_oracle.group_local_da_loss and pinned to the reference's own methods by the golden vectors):
    fg/bg split by softmax(cls)[0] >= 0.5; DAF: groups of more than k=20 RoIs are replaced by the k-means object's
    centroids, which the reference never updates (= 10 random normal vectors, `draw_centroids(dim, device)`), smaller
    groups are padded with their top-scoring member; the source group is used alone when non-empty, else the target
    group (the `!=0 & ... !=0` test is always False); the InstanceAlignmentHead sees every RoI as a one-token sequence
    (DAF/MAF), the FC-only head the whole batch (Deep); FocalLoss (DAF) or CrossEntropy (MAF/Deep) on the sigmoid
    outputs; no gradient (the reference returns .item()).  Returns a 0-dim fp32 tensor instead of a Python float (one
    sync per group for the data-dependent shapes instead of one per RoI).  Heads in train mode apply dropout, as there."""
    if flavour not in ("daf", "maf", "deep"):
        raise ValueError("flavour must be 'daf', 'maf' or 'deep'")
    draw = draw_centroids or _draw_centroids
    dev = bbox_feats[0].device
    groups = {}
    for d in (0, 1):                                   # source, then target: the order of the reference's RNG draws
        p = torch.softmax(bbox_cls[d].float(), dim=-1)
        fg = p[:, 0] >= 0.5
        for name, m, score in (("fg", fg, p[:, 0]), ("bg", ~fg, p[:, 1])):
            f = bbox_feats[d][m]
            n = f.shape[0]
            if n and flavour == "daf":
                if n > k:
                    f = draw(f.shape[1], dev).to(f.dtype)
                elif n < k:
                    top = torch.argmax(torch.softmax(score[m], dim=-1), dim=0)
                    f = torch.cat([f, f[top].unsqueeze(0).expand(k - n, -1)], 0)
            groups[(name, d)] = f
    total = torch.zeros((), dtype=torch.float32, device=dev)
    for name, head in (("fg", head_fore), ("bg", head_back)):
        src, tar = groups[(name, 0)], groups[(name, 1)]
        if src.shape[0]:
            f, label = src, 0
        elif tar.shape[0]:
            f, label = tar, 1
        else:
            continue
        z = head.forward_logits(f.contiguous()) if flavour == "deep" else head.forward_logits(f.contiguous(), single_token=True)
        labels = torch.full((f.shape[0],), label, dtype=torch.long, device=dev)
        if flavour == "daf":
            total = total + F_.sigmoid_focal_loss2(torch.sigmoid(z), labels, 2.0, 0.25)
        else:
            total = total + F_.ce2(z, labels, True)[0]
    return total.detach()


class FocalLoss(nn.Module):
    """mmdet FocalLoss (use_sigmoid=True) for the DA path: [k,2] predictions, integer targets."""

    def __init__(self, use_sigmoid=True, gamma=2.0, alpha=0.25, reduction="mean", loss_weight=1.0):
        super().__init__()
        if not use_sigmoid:
            raise NotImplementedError("Only sigmoid focal loss supported now.")
        self.use_sigmoid = use_sigmoid
        self.gamma, self.alpha, self.reduction, self.loss_weight = gamma, alpha, reduction, loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None):
        if weight is not None or avg_factor is not None or (reduction_override or self.reduction) != "mean":
            raise NotImplementedError("DA path uses mean reduction without weights")
        return self.loss_weight * F_.sigmoid_focal_loss2(pred, target, self.gamma, self.alpha)
