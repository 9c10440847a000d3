"""DA detectors with the reference's registry names, constructor signature and losses-dict keys:

  DAFasterRCNN_Org   mmdet/models/detectors/DAFaster_rcnn_Orig.py:11-229  local_da_loss, globle_da_loss, consistency_loss
  DAFasterRCNN       mmdet/models/detectors/DAFaster_rcnn.py:11-381       globle_da_loss, patch_bottom_loss, local_da_loss
  MAFasterRCNN       mmdet/models/detectors/MAFaster_rcnn.py:11-353       globle_da_loss, local_da_loss
  DAFasterRCNN_Deep  mmdet/models/detectors/DAFaster_rcnn_Deep.py:11-383  globle_da_loss, patch_bottom_loss, local_da_loss

Instance loss of DAF/MAF/Deep: the reference computes it per RoI in Python loops, DETACHES it (`.item()`, Q2),
classifies only the source RoIs (Q3) and, for DAF, first replaces the features by k-means centroids with a
random initialisation (non-deterministic).  Here it is one batched pass over all RoIs of both domains through the
fore/back heads (selected by softmax(cls)[0] >= 0.5) with the reference's criterion (focal for DAF, CE for
MAF/Deep) and `detach_instance_loss=True` reproduces the reference's "logged but gradient-free" behaviour."""
import torch
import torch.nn as nn

from . import da_heads, da_losses
from .hotpath import domain_tensor, parse_losses
from .registry import DETECTORS, build_backbone, build_head


class _DATwoStage(nn.Module):
    def __init__(self, backbone, rpn_head, roi_head, train_cfg, test_cfg, neck=None, pretrained=None, init_cfg=None):
        super().__init__()
        if neck is not None:
            raise NotImplementedError("no DA config of the reference uses a neck (SURVEY.md §2.4)")
        self.backbone = build_backbone(backbone)
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        rpn = dict(rpn_head)
        rpn.update(train_cfg=(train_cfg or {}).get("rpn"), test_cfg=(test_cfg or {}).get("rpn"))
        self.rpn_head = build_head(rpn)
        roi = dict(roi_head)
        roi.update(train_cfg=(train_cfg or {}).get("rcnn"), test_cfg=(test_cfg or {}).get("rcnn"))
        self.roi_head = build_head(roi)
        self.criterion = nn.CrossEntropyLoss()
        self.global_align = True
        self.local_align = True

    with_neck = False
    with_rpn = True

    def extract_feat(self, img):
        return self.backbone(img)

    def unused_parameters(self):
        out = self.backbone.unused_parameters() if hasattr(self.backbone, "unused_parameters") else []
        for m in self.children():
            if isinstance(m, (da_heads.InstanceAlignmentHead,)):
                out += m.unused_parameters()
        return out

    def _rpn_and_roi(self, x, img_metas, gt_bboxes, gt_labels, gt_da, gt_domain, gt_bboxes_ignore, gt_masks, proposals, kwargs):
        losses = dict()
        if proposals is None:
            proposal_cfg = (self.train_cfg or {}).get("rpn_proposal", (self.test_cfg or {}).get("rpn"))
            rpn_losses, proposal_list = self.rpn_head.forward_train(x, img_metas, gt_bboxes, gt_da=gt_domain, gt_labels=None,
                                                                    gt_bboxes_ignore=gt_bboxes_ignore, proposal_cfg=proposal_cfg)
            if rpn_losses is None:
                z = x[0].new_zeros(())
                rpn_losses = dict(loss_rpn_cls=z, loss_rpn_bbox=z)
            losses.update(rpn_losses)
        else:
            proposal_list = proposals
        roi_losses, bbox_feats, bbox_cls = self.roi_head.forward_train(x, img_metas, proposal_list, gt_bboxes, gt_labels, gt_da,
                                                                       gt_bboxes_ignore, gt_masks, **kwargs)
        losses.update(roi_losses)
        return losses, bbox_feats, bbox_cls

    def train_step(self, data, optimizer=None):
        losses = self.forward_train(**data)
        loss, log_vars = parse_losses(losses)
        return dict(loss=loss, log_vars=log_vars, num_samples=len(data["img_metas"]))

    def forward(self, img, img_metas, return_loss=True, **kwargs):
        if not return_loss:
            raise NotImplementedError("inference runs no DA module (SURVEY.md §3.4) and is out of scope")
        return self.forward_train(img, img_metas, **kwargs)


@DETECTORS.register_module()
class DAFasterRCNN_Org(_DATwoStage):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.local_da = da_heads.InstanceAlignmentHead()
        self.local_da._init_weights()

    def forward_train(self, img, img_metas, gt_bboxes, gt_labels, gt_da=None, gt_bboxes_ignore=None, gt_masks=None,
                      proposals=None, **kwargs):
        gt_domain = domain_tensor([int(d) for d in gt_da], img.device)
        x, global_loss, imgs_feat = self.backbone.forward_train(img, gt_domain)
        losses, bbox_feats, _ = self._rpn_and_roi(x, img_metas, gt_bboxes, gt_labels, gt_da, gt_domain, gt_bboxes_ignore,
                                                  gt_masks, proposals, kwargs)
        if self.local_align:
            local_da_loss, ins_preds, ins_labels = self.local_da_loss(bbox_feats, 0.1)
            losses.update(local_da_loss=0.1 * local_da_loss)
        if self.global_align:
            losses.update(globle_da_loss=0.1 * global_loss)
        losses.update(consistency_loss=0.1 * self.consist_loss(imgs_feat, ins_preds, ins_labels))
        return losses

    def local_da_loss(self, bbox_feats, lamda):
        labels = torch.cat([torch.full((len(f),), i, dtype=torch.long, device=f.device) for i, f in enumerate(bbox_feats)])
        loss, pred = self.local_da.forward_loss(torch.cat(list(bbox_feats), 0), labels)
        return loss, pred, labels

    def consist_loss(self, imgs_feat, ins_preds, ins_labels):
        return da_losses.consistency_loss(imgs_feat, ins_preds, ins_labels)


class _ForeBack(_DATwoStage):
    head_cls = da_heads.InstanceAlignmentHead
    group_flavour = "maf"
    use_focal = False
    local_lamda = 0.1
    has_patch = True
    detach_instance_loss = True   # Q2: the reference adds a Python float

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.criterion_fl = da_losses.FocalLoss(use_sigmoid=True, gamma=2.0, alpha=0.25, reduction="mean")
        self.local_da_fore = self.head_cls()
        self.local_da_back = self.head_cls()
        self.local_da_fore._init_weights()
        self.local_da_back._init_weights()

    def unused_parameters(self):
        """The fore/back heads are only ever evaluated without gradient (Q2: the reference adds their loss as a Python
        float), so EVERY parameter of both heads stays without .grad on every rank and every step: listed statically so
        the data-parallel reducer and the optimizer skip them (the reference's torch.optim.SGD skips grad=None too)."""
        out = super().unused_parameters()
        if self.detach_instance_loss:
            seen = {id(p) for p in out}
            out += [p for h in (self.local_da_fore, self.local_da_back) for p in h.parameters() if id(p) not in seen]
        return out

    def group_local_da_loss(self, bbox_feats, lamda, bbox_cls):
        """L5 as the reference computes it (da_losses.group_local_da_loss): detached, source group first."""
        return da_losses.group_local_da_loss(bbox_feats, bbox_cls, self.local_da_fore, self.local_da_back, self.group_flavour)

    def forward_train(self, img, img_metas, gt_bboxes, gt_labels, gt_da=None, gt_bboxes_ignore=None, gt_masks=None,
                      proposals=None, **kwargs):
        gt_domain = domain_tensor([int(d) for d in gt_da], img.device)
        out = self.backbone.forward_train(img, gt_domain)
        x, global_loss = out[0], out[1]
        losses, bbox_feats, bbox_cls = self._rpn_and_roi(x, img_metas, gt_bboxes, gt_labels, gt_da, gt_domain, gt_bboxes_ignore,
                                                         gt_masks, proposals, kwargs)
        if self.local_align:
            losses.update(local_da_loss=self.local_lamda * self.group_local_da_loss(bbox_feats, self.local_lamda, bbox_cls))
        if self.global_align:
            losses.update(globle_da_loss=0.1 * global_loss.sum())
            if self.has_patch:
                losses.update(patch_bottom_loss=0.1 * out[2])
        return losses


@DETECTORS.register_module()
class DAFasterRCNN(_ForeBack):
    use_focal, local_lamda, group_flavour = True, 0.2, "daf"


@DETECTORS.register_module()
class MAFasterRCNN(_ForeBack):
    has_patch = False


@DETECTORS.register_module()
class DAFasterRCNN_Deep(_ForeBack):
    head_cls = da_heads.InstanceAlignmentHead_DAF
    group_flavour = "deep"
    local_lamda = 0.2      # DAFaster_rcnn_Deep.py:177
