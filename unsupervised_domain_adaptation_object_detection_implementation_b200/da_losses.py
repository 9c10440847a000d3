"""Domain losses of the DA path as the reference composes them.

  daf_image_loss          L1  mmdet/models/backbones/resnet_da_daf_org.py:816-822
  patch_loss              L2  mmdet/models/backbones/resnet_da_cbam.py:971-979
  image_ce_loss           L3  resnet_da_cbam.py:966-968 (raw logits), resnet_da.py:846-848 (on sigmoid)
  instance_ce_loss        L4  mmdet/models/detectors/DAFaster_rcnn_Orig.py:177-188
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
