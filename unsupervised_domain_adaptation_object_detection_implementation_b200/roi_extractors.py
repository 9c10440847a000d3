"""RoI construction and extraction on the DA path.

  bbox2roi / bbox2roi_train   mmdet/core/bbox/transforms.py:59-78,
                              mmdet/models/roi_heads/standard_roi_head_da_v5.py:12-33   (R1)
  SingleRoIExtractor          mmdet/models/roi_heads/roi_extractors/single_level_roi_extractor.py:9-115,
                              base_roi_extractor.py:37-60                               (R2)
"""
import torch
import torch.nn as nn

from . import functional as F_
from . import ops


_IDX_COL = {}


def bbox2roi(bbox_list):
    """list of [n_i,4+] boxes -> [sum n_i, 5] rois (batch_ind, x1, y1, x2, y2); batch_ind is the
    image index stored as float, exactly as the reference builds it.  Same values as the reference's per-image
    new_full + cat + cat (5 launches for a pair); here the index column is a cached constant of (sizes, dtype, device) and
    the boxes are written straight into their 4 columns: one copy per image."""
    if len(bbox_list) == 0:
        return torch.zeros((0, 5))
    sizes = tuple(int(b.size(0)) for b in bbox_list)
    ref = bbox_list[0]
    key = (sizes, ref.dtype, str(ref.device))
    col = _IDX_COL.get(key)
    if col is None:
        col = torch.cat([ref.new_full((n, 1), i) for i, n in enumerate(sizes)]) if sum(sizes) else ref.new_zeros((0, 1))
        _IDX_COL[key] = col
    rois = ref.new_empty((sum(sizes), 5))
    rois[:, :1] = col
    start = 0
    for b, n in zip(bbox_list, sizes):
        if n:
            rois[start:start + n, 1:] = b[:, :4]
        start += n
    return rois


def bbox2roi_train(bbox_list):
    """Per-image list variant (standard_roi_head_da_v5.py:12-33).  batch_ind = image index; the
    caller must pool from the FULL [N,C,H,W] map (the reference slices the map to
    batch size 1 and reads out of bounds for image 1)."""
    rois = bbox2roi(bbox_list)
    out, start = [], 0
    for b in bbox_list:
        out.append(rois[start:start + b.size(0)])
        start += b.size(0)
    return out


class SingleRoIExtractor(nn.Module):
    """Same config surface as the reference: roi_layer=dict(type='RoIAlign', output_size=7,
    sampling_ratio=0), out_channels, featmap_strides, finest_scale."""

    def __init__(self, roi_layer, out_channels, featmap_strides, finest_scale=56, init_cfg=None):
        super().__init__()
        cfg = dict(roi_layer)
        layer_type = cfg.pop("type")
        layer_cls = getattr(ops, layer_type)  # looked up by name like mmcv.ops
        self.roi_layers = nn.ModuleList([layer_cls(spatial_scale=1.0 / s, **cfg) for s in featmap_strides])
        self.out_channels = out_channels
        self.featmap_strides = list(featmap_strides)
        self.finest_scale = finest_scale
        self.fp16_enabled = False

    @property
    def num_inputs(self):
        return len(self.featmap_strides)

    def map_roi_levels(self, rois, num_levels):
        return F_.map_roi_levels(rois, num_levels, float(self.finest_scale))

    def forward(self, feats, rois, roi_scale_factor=None):
        if roi_scale_factor is not None:
            raise NotImplementedError("roi_scale_factor is not used on the DA path")
        out_size = self.roi_layers[0].output_size
        num_levels = len(feats)
        if num_levels == 1:
            if len(rois) == 0:
                return feats[0].new_zeros(0, self.out_channels, *out_size)
            return self.roi_layers[0](feats[0], rois)
        if len(rois) == 0:
            return feats[0].new_zeros(0, self.out_channels, *out_size)
        # FPN path (single_level_roi_extractor.py:81-104).  The reference gathers the RoIs of each level with nonzero() -- a
        # host sync per level, and not capturable in a CUDA graph.  Here every level sees the WHOLE RoI list with the batch
        # index of the RoIs that belong to another level set to -1: the kernels treat such a RoI as empty (zero footprint,
        # zeros written, no feature read), so the per-level outputs have disjoint supports and their sum is the reference's
        # scatter -- same values, no host round trip, fixed launch sequence.
        target_lvls = self.map_roi_levels(rois, num_levels)
        roi_feats = None
        for i in range(num_levels):
            masked = rois.clone()
            masked[:, 0] = torch.where(target_lvls == i, rois[:, 0], rois.new_full((), -1.0))
            out_i = self.roi_layers[i](feats[i], masked)
            roi_feats = out_i if roi_feats is None else roi_feats + out_i
        return roi_feats
