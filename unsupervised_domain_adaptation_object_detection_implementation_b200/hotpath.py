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

    def __init__(self, in_channels=2048, roi_feat_size=7, fc_out_channels=1024, num_shared_fcs=2, roi_layout="rchw"):
        """roi_layout="rhwc": the RoI features arrive bin-major ([R,7,7,C] memory, logical [R,C,7,7]) and the first FC's
        weight is HELD with its input columns in (bin, channel) order, so that `flatten` is a view and the tensor-core RoIAlign
        backward gets its operand as it lies in memory.  state_dict()/load_state_dict() still speak the
        reference's (channel, bin) column order (convfc_bbox_head.py:229: `x.flatten(1)` of [R,C,7,7]); the optimizer is
        elementwise, so training in the permuted order is the same training.  Optimizer STATE of shared_fcs[0].weight (momentum
        in FusedSGD.state_dict()) is in the held order: a checkpoint resumes with the same roi_layout, or its buffer goes
        through to_reference_order / to_held_order like the weight."""
        super().__init__()
        if roi_layout not in ("rchw", "rhwc"):
            raise ValueError(f"roi_layout {roi_layout!r}")
        self.roi_layout = roi_layout
        self.in_channels, self.bins = in_channels, roi_feat_size * roi_feat_size
        dims = [in_channels * roi_feat_size * roi_feat_size] + [fc_out_channels] * num_shared_fcs
        self.shared_fcs = nn.ModuleList([nn.Linear(dims[i], dims[i + 1]) for i in range(num_shared_fcs)])
        for fc in self.shared_fcs:
            nn.init.xavier_uniform_(fc.weight)
            nn.init.constant_(fc.bias, 0)
        if roi_layout == "rhwc":
            with torch.no_grad():
                self.shared_fcs[0].weight.copy_(self.to_held_order(self.shared_fcs[0].weight))
            self._register_state_dict_hook(SharedFCs._export_reference_order)
            self._register_load_state_dict_pre_hook(self._import_reference_order)

    # ---- (channel, bin) <-> (bin, channel) column order of the first FC (roi_layout="rhwc" only)
    def to_held_order(self, w):
        """[out, C*49] in the reference's (channel, bin) column order -> the order this module holds."""
        if self.roi_layout != "rhwc":
            return w
        return w.reshape(w.shape[0], self.in_channels, self.bins).transpose(1, 2).reshape(w.shape[0], -1)

    def to_reference_order(self, w):
        """Inverse of to_held_order (weights, gradients and momentum buffers of shared_fcs[0].weight alike)."""
        if self.roi_layout != "rhwc":
            return w
        return w.reshape(w.shape[0], self.bins, self.in_channels).transpose(1, 2).reshape(w.shape[0], -1)

    @staticmethod
    def _export_reference_order(module, state_dict, prefix, local_metadata):
        key = prefix + "shared_fcs.0.weight"
        if key in state_dict:
            state_dict[key] = module.to_reference_order(state_dict[key])

    def _import_reference_order(self, state_dict, prefix, *args):
        key = prefix + "shared_fcs.0.weight"
        if key in state_dict and state_dict[key].dim() == 2 and state_dict[key].shape[1] == self.in_channels * self.bins:
            state_dict[key] = self.to_held_order(state_dict[key])

    def _flatten(self, x):
        if self.roi_layout == "rhwc":
            if x.dim() != 4:
                raise RuntimeError("SharedFCs(roi_layout='rhwc') takes the [R,C,7,7] RoI features, not a flattened tensor")
            x = x.permute(0, 2, 3, 1)                     # a view when the RoI tensor is bin-major; else one re-layout pass
            return x.reshape(x.shape[0], -1)
        return x.flatten(1)

    def forward(self, x):
        x = F_.cast(self._flatten(x), F_.act_dtype())
        k = x.shape[0]
        x = x.view(k, 1, 1, -1)
        for fc in self.shared_fcs:
            x = F_.dense_layer(x, fc.weight, None, fc.bias, relu=True)
        return x.view(k, -1)

    def can_feed_chain(self, x):
        fcs = self.shared_fcs
        return (len(fcs) >= 2 and da_heads.USE_CHAIN and da_heads.USE_CHAIN_FEED and F_.get_engine() == "umma_bf16"
                and torch.is_grad_enabled() and x.is_cuda and x.shape[0] > 0
                and fcs[-1].in_features % 64 == 0 and fcs[-1].out_features % 64 == 0)

    def split(self, x):
        """All layers but the last, and the last one as the `pre` tuple of InstanceAlignmentHead*.forward_loss: the instance
        head's chain kernel runs it (forward and backward), applies the ReLU derivative of the layer before it and returns
        that layer's bias gradient, so the layer before runs with preact_grad=True (no activation-backward / bias-sum kernels).
        Only for callers whose ONLY consumer of the shared features is the instance head (the hot path)."""
        x = F_.cast(self._flatten(x), F_.act_dtype())
        k = x.shape[0]
        x = x.view(k, 1, 1, -1)
        fcs = list(self.shared_fcs)
        for i, fc in enumerate(fcs[:-1]):
            last = i == len(fcs) - 2
            x = F_.dense_layer(x, fc.weight, None, fc.bias.detach() if last else fc.bias, relu=True, preact_grad=last)
        b_in = fcs[-2].bias if len(fcs) >= 2 else None
        return x.view(k, -1), (fcs[-1].weight, fcs[-1].bias, b_in)


def instance_branch(bbox_head, local_da, roi_feats, label_da):
    """RoI features -> shared FCs -> instance head + its CE loss.  When the shared features have no other consumer and the
    shapes allow it, the LAST shared FC runs inside the instance head's chain kernel (one kernel per direction)."""
    if bbox_head is not None and bbox_head.can_feed_chain(roi_feats):
        xin, (w0, b0, b_in) = bbox_head.split(roi_feats)
        return local_da.forward_loss(None, label_da, pre=(xin, w0, b0, b_in))
    bbox_feats = bbox_head(roi_feats) if bbox_head is not None else roi_feats.flatten(1)
    return local_da.forward_loss(bbox_feats, label_da)


class DAFOrgHotPath(nn.Module):
    def __init__(self, in_channels=2048, featmap_stride=16, fc_out_channels=1024,
                 lambdas=(0.1, 0.1, 0.1), with_shared_fcs=True, roi_layout="rchw"):
        """roi_layout="rhwc": RoI features bin-major between RoIAlign and the first shared FC (see SharedFCs); same losses and
        gradients up to the summation order of that FC's 100352-long dot products."""
        super().__init__()
        if roi_layout == "rhwc" and not with_shared_fcs:
            raise ValueError("roi_layout='rhwc' needs the shared FCs (they hold the permuted weight)")
        self.da_head_top = da_heads.ImgAlignmentHead(in_channels)
        self.da_head_top._init_weights()
        self.bbox_roi_extractor = SingleRoIExtractor(dict(type="RoIAlign", output_size=7, sampling_ratio=0),
                                                     out_channels=in_channels, featmap_strides=[featmap_stride])
        for layer in self.bbox_roi_extractor.roi_layers:
            layer.out_layout = roi_layout
        self.bbox_head = SharedFCs(in_channels, 7, fc_out_channels, roi_layout=roi_layout) if with_shared_fcs else None
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
        label_da = roi_domain_labels(tuple(len(p) for p in proposal_list), c5.device) if len(proposal_list) == 2 else \
            rois[:, 0].to(torch.int32).clamp(max=1)
        ins_loss, ins_preds = instance_branch(self.bbox_head, self.local_da, roi_feats, label_da)
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


class DeepHotPath(nn.Module):
    """Image-level + instance-level part of DAFasterRCNN_Deep / ResNet_DA_Deep (detectors/DAFaster_rcnn_Deep.py:78-196,
    backbones/resnet_da_deep.py:1120-1175): Global heads without the dead branch on C4/C5 (CE on raw logits, L3),
    NonLocalAlignmentHead on C3 (and C4 when `patch_mid`) with the per-image patch loss L2 over all C*H*W values, and the
    FC-only InstanceAlignmentHead_DAF on the RoI features (one chain kernel).  The T x T attention of the NonLocalBlock is
    materialised (T = H*W of the level): fine at the reference's 512x1024 scale (C3: T = 8192), not at 1024x2048
    (T = 32768 -> 4.3 GB per image; the blocked two-pass kernel is SURVEY 8f rank 2, not built)."""

    def __init__(self, channels=(512, 1024, 2048), fc_out_channels=1024, lambdas=(0.1, 0.1, 0.2), patch_mid=False):
        super().__init__()
        self.da_head_mid = da_heads.GlobalAlignmentHeadDeep(channels[1])
        self.da_head_top = da_heads.GlobalAlignmentHeadDeep(channels[2])
        self.local_da_head_bottom = da_heads.NonLocalAlignmentHead(channels[0])
        self.local_da_head_mid = da_heads.NonLocalAlignmentHead(channels[1]) if patch_mid else None
        self.da_head_mid._init_weights()
        self.da_head_top._init_weights()
        self.local_da = da_heads.InstanceAlignmentHead_DAF()
        self.local_da._init_weights()
        self.global_lamda, self.patch_lamda, self.local_lamda = lambdas

    def unused_parameters(self):
        return []

    def forward_train(self, c3, c4, c5, bbox_feats, gt_da):
        """bbox_feats: [R_src + R_tgt, fc_out] features of the shared FCs (source RoIs first); -> losses dict."""
        gt_domain = domain_tensor(gt_da, c5.device)
        g_mid, _ = da_losses.image_ce_loss(self.da_head_mid(c4), gt_domain, False)
        g_top, _ = da_losses.image_ce_loss(self.da_head_top(c5), gt_domain, False)
        patch = da_losses.patch_loss(self.local_da_head_bottom(c3), gt_domain)
        if self.local_da_head_mid is not None:
            patch = patch + da_losses.patch_loss(self.local_da_head_mid(c4), gt_domain)
        half = bbox_feats.shape[0] // 2
        labels = roi_domain_labels((half, bbox_feats.shape[0] - half), c5.device)
        ins_loss, _ = self.local_da.forward_loss(bbox_feats, labels)
        return dict(globle_da_loss=self.global_lamda * (g_mid + g_top), patch_bottom_loss=self.patch_lamda * patch,
                    local_da_loss=self.local_lamda * ins_loss)


class FPNHotPath(nn.Module):
    """The DAF hot path on FPN levels (BASELINE config 2b; SURVEY 8d: an EXTENSION -- no DA config of the reference has a
    neck).  One ImgAlignmentHead per level with its per-level mean loss (fused tail kernel), RoIAlign through the multi-level
    SingleRoIExtractor (FPN level map, single_level_roi_extractor.py:36-55; device-side partition, CUDA-graph capturable),
    shared FCs (channels*49 -> fc_out -> fc_out), InstanceAlignmentHead + CE (chain kernel), consistency against the
    stride-16 level."""

    def __init__(self, channels=256, featmap_strides=(4, 8, 16, 32), fc_out_channels=1024, lambdas=(0.1, 0.1, 0.1)):
        super().__init__()
        self.featmap_strides = list(featmap_strides)
        self.da_heads = nn.ModuleList([da_heads.ImgAlignmentHead(channels) for _ in self.featmap_strides])
        for h in self.da_heads:
            h._init_weights()
        self.bbox_roi_extractor = SingleRoIExtractor(dict(type="RoIAlign", output_size=7, sampling_ratio=0),
                                                     out_channels=channels, featmap_strides=self.featmap_strides)
        self.bbox_head = SharedFCs(channels, 7, fc_out_channels)
        self.local_da = da_heads.InstanceAlignmentHead()
        self.local_da._init_weights()
        self.global_lamda, self.local_lamda, self.consist_lamda = lambdas
        self.consist_level = self.featmap_strides.index(16) if 16 in self.featmap_strides else len(self.featmap_strides) - 1

    def unused_parameters(self):
        return self.local_da.unused_parameters()

    def forward_train(self, feats, proposal_list, gt_da):
        gt_domain = domain_tensor(gt_da, feats[0].device)
        level_losses, level_feats = zip(*[h.forward_loss(f, gt_domain) for h, f in zip(self.da_heads, feats)])
        rois = bbox2roi(proposal_list)
        roi_feats = self.bbox_roi_extractor(list(feats), rois)
        label_da = roi_domain_labels(tuple(len(p) for p in proposal_list), feats[0].device) if len(proposal_list) == 2 else \
            rois[:, 0].to(torch.int32).clamp(max=1)
        ins_loss, ins_preds = instance_branch(self.bbox_head, self.local_da, roi_feats, label_da)
        consist = da_losses.consistency_loss(level_feats[self.consist_level], ins_preds, label_da)
        n = len(level_losses)
        scaled, total = F_.weighted_losses(list(level_losses) + [ins_loss, consist],
                                           [self.global_lamda] * n + [self.local_lamda, self.consist_lamda])
        out = LossDict(total, local_da_loss=scaled[n], consistency_loss=scaled[n + 1])
        out["globle_da_loss"] = [scaled[i] for i in range(n)]     # one entry per level; _parse_losses sums a list (base.py:196-197)
        return out
