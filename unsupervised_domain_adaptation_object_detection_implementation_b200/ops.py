"""`ops` namespace mirroring the slice of `mmcv.ops` the DA path looks up by name
(mmdet/models/roi_heads/roi_extractors/base_roi_extractor.py:54-60 does
`getattr(ops, 'RoIAlign')(spatial_scale=1/s, output_size=7, sampling_ratio=0)`;
mmdet/models/losses/focal_loss.py:5 imports `sigmoid_focal_loss`)."""
import torch
import torch.nn as nn
from torch.nn.modules.utils import _pair

from . import functional as F_


class RoIAlign(nn.Module):
    """Same constructor and call signature as mmcv.ops.RoIAlign (mmcv-full 1.3.17).

    pool_mode: only 'avg' (the DA configs never use 'max').
    use_torchvision: accepted for signature compatibility; ignored (no library fallback).
    """

    def __init__(self, output_size, spatial_scale=1.0, sampling_ratio=0, pool_mode="avg", aligned=True,
                 use_torchvision=False):
        super().__init__()
        if pool_mode != "avg":
            raise NotImplementedError("libda_b200 RoIAlign implements pool_mode='avg' only")
        self.output_size = _pair(output_size)
        self.spatial_scale = float(spatial_scale)
        self.sampling_ratio = int(sampling_ratio)
        self.pool_mode = pool_mode
        self.aligned = aligned
        self.use_torchvision = use_torchvision
        # "rhwc": the SAME logical [R,C,7,7] result, stored bin-major ([R,7,7,C] memory = torch.channels_last).  Not part of
        # mmcv's signature: set by callers that consume channels-last RoI features (hotpath.SharedFCs, roi_layout="rhwc").
        self.out_layout = "rchw"

    def forward(self, input, rois):
        """input: NCHW feature map; rois: [R,5] (batch_index, x1, y1, x2, y2)."""
        if self.out_layout == "rhwc":
            return F_.roi_align(input, rois, self.output_size, self.spatial_scale, self.sampling_ratio, self.aligned,
                                out_layout="rhwc").permute(0, 3, 1, 2)
        return F_.roi_align(input, rois, self.output_size, self.spatial_scale, self.sampling_ratio, self.aligned)

    def __repr__(self):
        return (f"{self.__class__.__name__}(output_size={self.output_size}, spatial_scale={self.spatial_scale}, "
                f"sampling_ratio={self.sampling_ratio}, pool_mode={self.pool_mode}, aligned={self.aligned}, "
                f"use_torchvision={self.use_torchvision})")


def roi_align(input, rois, output_size, spatial_scale=1.0, sampling_ratio=0, pool_mode="avg", aligned=True):
    if pool_mode != "avg":
        raise NotImplementedError("libda_b200 roi_align implements pool_mode='avg' only")
    return F_.roi_align(input, rois, _pair(output_size), spatial_scale, sampling_ratio, aligned)


def sigmoid_focal_loss(pred, target, gamma=2.0, alpha=0.25, weight=None, reduction="mean"):
    """mmcv.ops.sigmoid_focal_loss for the DA path's [k,2] predictions (mean reduction)."""
    if weight is not None or reduction != "mean" or pred.dim() != 2 or pred.shape[1] != 2:
        raise NotImplementedError("sigmoid_focal_loss: the DA path uses [k,2] inputs, no weight, mean reduction")
    return F_.sigmoid_focal_loss2(pred, target, gamma, alpha)
