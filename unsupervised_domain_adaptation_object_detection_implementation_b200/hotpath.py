"""The DA hot path as one module per reference detector flavour (what `forward_train` adds on top
of a plain Faster R-CNN), operating on backbone features and sampled RoIs.

  DAFOrgHotPath   mmdet/models/detectors/DAFaster_rcnn_Orig.py:59-188 +
                  mmdet/models/backbones/resnet_da_daf_org.py:796-824            (H1,L1,R2,F1,I1,L4,L7)
  CBAMHotPath     mmdet/models/detectors/DAFaster_rcnn.py:78-196 +
                  mmdet/models/backbones/resnet_da_cbam.py:934-993               (H2,L2,H3x2,L3)
  MAFHotPath      mmdet/models/detectors/MAFaster_rcnn.py:66-168 +
                  mmdet/models/backbones/resnet_da.py:821-850                    (H4x3,L3)

Loss-dict keys and lambda weights are the reference's (W1).  RoIs are pooled from the FULL [N,C,H,W]
map with batch_ind = image index (the intended semantics of standard_roi_head_da_v5.py:199-217, Q1).
"""
import torch
import torch.nn as nn

from . import da_heads, da_losses, functional as F_
from .roi_extractors import SingleRoIExtractor, bbox2roi


_CONST = {}


def domain_tensor(gt_da, device):
    """Device copy of the per-image domain labels, cached per (labels, device): building it every step
    would be a pageable host->device copy (a sync the reference pays at DAFaster_rcnn_Orig.py:119-122, and
    illegal inside CUDA-graph capture)."""
    if torch.is_tensor(gt_da):
        return gt_da.to(device=device, dtype=torch.int32)
    key = (tuple(int(d) for d in gt_da), str(device))
    t = _CONST.get(key)
    if t is None:
        t = torch.tensor(key[0], dtype=torch.int32, device=device)      # int32: what the kernels read (no per-step cast)
        _CONST[key] = t
    return t


def roi_domain_labels(counts, device):
    """Per-RoI domain labels (image 0 = source, image 1 = target), cached per (counts, device): a constant of the
    step, not three fill/cat kernels per iteration."""
    key = ("roi_labels", counts, str(device))
    t = _CONST.get(key)
    if t is None:
        t = torch.cat([torch.full((n,), d, dtype=torch.int32, device=device) for n, d in zip(counts, (0, 1))])
        _CONST[key] = t
    return t


class LossDict(dict):
    """The reference's losses dict (same keys) that also carries the already computed sum of its entries."""

    def __init__(self, total, **entries):
        super().__init__(**entries)
        self.total = total


class SharedFCs(nn.Module):
    """Shared2FCBBoxHead's shared_fcs as used by forward_train_da
    (mmdet/models/roi_heads/bbox_heads/convfc_bbox_head.py:198-237): flatten(1) -> FC -> ReLU -> FC -> ReLU.
    Adjacent to the hot path; it produces the 1024-d features the instance head consumes."""

    def __init__(self, in_channels=2048, roi_feat_size=7, fc_out_channels=1024, num_shared_fcs=2):
        super().__init__()
        dims = [in_channels * roi_feat_size * roi_feat_size] + [fc_out_channels] * num_shared_fcs
        self.shared_fcs = nn.ModuleList([nn.Linear(dims[i], dims[i + 1]) for i in range(num_shared_fcs)])
        for fc in self.shared_fcs:
            nn.init.xavier_uniform_(fc.weight)
            nn.init.constant_(fc.bias, 0)

    def forward(self, x):
        x = F_.cast(x.flatten(1), F_.act_dtype())
        k = x.shape[0]
        x = x.view(k, 1, 1, -1)
        for fc in self.shared_fcs:
            x = F_.dense_layer(x, fc.weight, None, fc.bias, relu=True)
        return x.view(k, -1)


class DAFOrgHotPath(nn.Module):
    def __init__(self, in_channels=2048, featmap_stride=16, fc_out_channels=1024,
                 lambdas=(0.1, 0.1, 0.1), with_shared_fcs=True):
        super().__init__()
        self.da_head_top = da_heads.ImgAlignmentHead(in_channels)
        self.da_head_top._init_weights()
        self.bbox_roi_extractor = SingleRoIExtractor(dict(type="RoIAlign", output_size=7, sampling_ratio=0),
                                                     out_channels=in_channels, featmap_strides=[featmap_stride])
        self.bbox_head = SharedFCs(in_channels, 7, fc_out_channels) if with_shared_fcs else None
        self.local_da = da_heads.InstanceAlignmentHead()
        self.local_da._init_weights()
        self.global_lamda, self.local_lamda, self.consist_lamda = lambdas

    def unused_parameters(self):
        return self.local_da.unused_parameters()

    def forward_train(self, c5, proposal_list, gt_da):
        """c5 [N,C,H,W]; proposal_list: per-image [n_i,4] boxes (image 0 = source, 1 = target);
        gt_da: per-image domain (0/1).  Returns the reference's DA entries of the losses dict."""
        gt_domain = domain_tensor(gt_da, c5.device)
        global_loss, imgs_feat = self.da_head_top.forward_loss(c5, gt_domain)      # H1 + L1: GEMM + one fused tail kernel
        rois = bbox2roi(proposal_list)
        roi_feats = self.bbox_roi_extractor([c5], rois)
        bbox_feats = self.bbox_head(roi_feats) if self.bbox_head is not None else roi_feats.flatten(1)
        label_da = roi_domain_labels(tuple(len(p) for p in proposal_list), c5.device) if len(proposal_list) == 2 else \
            rois[:, 0].to(torch.int32).clamp(max=1)
        ins_loss, ins_preds = self.local_da.forward_loss(bbox_feats, label_da)
        consist = da_losses.consistency_loss(imgs_feat, ins_preds, label_da)
        # lambda weights (DAFaster_rcnn_Orig.py:143-157) and the total of the dict in ONE launch; parse_losses picks the total up
        scaled, total = F_.weighted_losses([ins_loss, global_loss, consist], [self.local_lamda, self.global_lamda, self.consist_lamda])
        return LossDict(total, local_da_loss=scaled[0], globle_da_loss=scaled[1], consistency_loss=scaled[2])


class CBAMHotPath(nn.Module):
    """Image-level part of DAFasterRCNN / ResNet_DA_CBAM: Local head on C3, Global heads on C4, C5."""

    def __init__(self, channels=(512, 1024, 2048), lambdas=(0.1, 0.1)):
        super().__init__()
        self.local_da_head_bottom = da_heads.LocalAlignmentHead(channels[0])
        self.da_head_mid = da_heads.GlobalAlignmentHead(channels[1])
        self.da_head_top = da_heads.GlobalAlignmentHead(channels[2])
        self.da_head_mid._init_weights()
        self.da_head_top._init_weights()
        self.global_lamda, self.patch_lamda = lambdas

    def unused_parameters(self):
        return self.da_head_mid.unused_parameters() + self.da_head_top.unused_parameters()

    def forward_train(self, c3, c4, c5, gt_da):
        gt_domain = domain_tensor(gt_da, c5.device)
        patch, local_feat = self.local_da_head_bottom.forward_loss(c3, gt_domain)   # H2 + L2: fused tail
        g_mid, _ = da_losses.image_ce_loss(self.da_head_mid(c4), gt_domain, False)
        g_top, _ = da_losses.image_ce_loss(self.da_head_top(c5), gt_domain, False)
        return dict(globle_da_loss=self.global_lamda * (g_mid + g_top), patch_bottom_loss=self.patch_lamda * patch)


class MAFHotPath(nn.Module):
    """Image-level part of MAFasterRCNN / ResNet_DA: SRM heads on C3, C4, C5 with CE on sigmoid outputs."""

    def __init__(self, channels=(512, 1024, 2048), global_lamda=0.1):
        super().__init__()
        self.da_head_bottom = da_heads.SRM(channels[0])
        self.da_head_mid = da_heads.SRM(channels[1])
        self.da_head_top = da_heads.SRM(channels[2])
        for h in (self.da_head_bottom, self.da_head_mid, self.da_head_top):
            h._init_weights()
        self.global_lamda = global_lamda

    def forward_train(self, c3, c4, c5, gt_da):
        gt_domain = domain_tensor(gt_da, c5.device)
        total = 0
        for head, feat in ((self.da_head_bottom, c3), (self.da_head_mid, c4), (self.da_head_top, c5)):
            loss, _ = da_losses.image_ce_loss(head.forward_logits(feat), gt_domain, True)
            total = total + loss
        return dict(globle_da_loss=self.global_lamda * total)


def parse_losses(losses):
    """mmdet/models/detectors/base.py:176-219: total = sum of every entry whose key contains 'loss'."""
    def _mean(v):   # the DA losses are device scalars already: no reduction kernel for a 0-d tensor
        return v if v.dim() == 0 else v.mean()
    log_vars = {k: _mean(v) if torch.is_tensor(v) else sum(_mean(x) for x in v) for k, v in losses.items()}
    if isinstance(losses, LossDict) and all("loss" in k for k in losses):
        loss = losses.total            # summed by the same kernel that applied the lambda weights
    else:
        loss = sum(v for k, v in log_vars.items() if "loss" in k)
    log_vars["loss"] = loss
    return loss, log_vars
