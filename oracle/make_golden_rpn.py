"""ORACLE — test infrastructure only.  Generates tests/golden/rpn_proposals.pt in the BUILD CONTAINER:
the reference's own AnchorGenerator (mmdet/core/anchor/anchor_generator.py) and delta2bbox
(mmdet/core/bbox/coder/delta_xywh_bbox_coder.py) are loaded IN PLACE from /root/reference under the mmcv stub of
oracle/ref_loader.py and run through the steps of RPNHeadDA._get_bboxes_single / _bbox_post_process
(mmdet/models/dense_heads/rpn_head_da.py:211-303); torchvision.ops.nms stands in for mmcv.ops.batched_nms (mmcv-full is absent;
one level => plain NMS).

    python -m oracle.make_golden_rpn
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader, seeded  # noqa: E402


def load_reference():
    ref_loader._install_stubs()
    reg = type("R", (), {"register_module": lambda self, *a, **k: (lambda c: c)})()
    for name in ["mmdet_ref.core", "mmdet_ref.core.bbox", "mmdet_ref.core.bbox.coder", "mmdet_ref.core.anchor"]:
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    b1 = types.ModuleType("mmdet_ref.core.bbox.builder")
    b1.BBOX_CODERS = reg
    sys.modules["mmdet_ref.core.bbox.builder"] = b1
    b2 = types.ModuleType("mmdet_ref.core.anchor.builder")
    b2.PRIOR_GENERATORS = reg
    sys.modules["mmdet_ref.core.anchor.builder"] = b2
    sys.modules["mmcv"].is_tuple_of = lambda seq, t: isinstance(seq, tuple) and all(isinstance(i, t) for i in seq)
    ref_loader._load("mmdet_ref.core.bbox.coder.base_bbox_coder", "mmdet/core/bbox/coder/base_bbox_coder.py")
    coder = ref_loader._load("mmdet_ref.core.bbox.coder.delta_xywh_bbox_coder", "mmdet/core/bbox/coder/delta_xywh_bbox_coder.py")
    anchor = ref_loader._load("mmdet_ref.core.anchor.anchor_generator", "mmdet/core/anchor/anchor_generator.py")
    return coder, anchor


def case(coder, anchor, name, H, W, nms_pre, max_per_img, img_shape, min_size=0, iou_thr=0.7, cls_scale=2.0, reg_scale=0.5):
    from torchvision.ops import nms
    gen = anchor.AnchorGenerator(strides=[16], ratios=[0.5, 1.0, 2.0], scales=[2, 4, 8, 16, 32])      # faster_rcnn_r50_torch_daf.py:26-30
    A = gen.num_base_anchors[0]
    cls = seeded.seeded_tensor(f"rpn.{name}.cls", (A, H, W), 0, scale=cls_scale)
    reg = seeded.seeded_tensor(f"rpn.{name}.reg", (4 * A, H, W), 0, scale=reg_scale)
    anchors = gen.grid_anchors([(H, W)], device="cpu")[0]
    scores = cls.permute(1, 2, 0).reshape(-1).sigmoid()                                               # rpn_head_da.py:226-229
    deltas = reg.permute(1, 2, 0).reshape(-1, 4)
    ranked, inds = scores.sort(descending=True, stable=True)                                          # :239
    n = nms_pre if 0 < nms_pre < scores.shape[0] else scores.shape[0]
    ranked, inds = ranked[:n], inds[:n]
    bbox_coder = coder.DeltaXYWHBBoxCoder(target_means=[.0, .0, .0, .0], target_stds=[1.0, 1.0, 1.0, 1.0])
    props = bbox_coder.decode(anchors[inds], deltas[inds], max_shape=img_shape)                       # :282-283
    decoded = props.clone()
    if min_size >= 0:                                                                                 # :290-297
        w, h = props[:, 2] - props[:, 0], props[:, 3] - props[:, 1]
        valid = (w > min_size) & (h > min_size)
        props, ranked = props[valid], ranked[valid]
    keep = nms(props, ranked, iou_thr)                                                                # batched_nms, one level
    dets = torch.cat([props[keep], ranked[keep, None]], -1)[:max_per_img]                             # :303
    return dict(H=H, W=W, A=A, stride=16, nms_pre=nms_pre, max_per_img=max_per_img, img_shape=tuple(img_shape), min_size=min_size,
                iou_thr=iou_thr, cls_scale=cls_scale, reg_scale=reg_scale, base_anchors=gen.base_anchors[0].clone(),
                anchors_first=anchors[:32].clone(), anchors_last=anchors[-32:].clone(), top_idx_head=inds[:64].clone(),
                decoded_head=decoded[:256].clone(), n_decoded=int(decoded.shape[0]), dets=dets.clone())


def main():
    coder, anchor = load_reference()
    out = {
        "small": case(coder, anchor, "small", 12, 20, 600, 100, (192, 320)),
        "min_size": case(coder, anchor, "min_size", 12, 20, 2000, 300, (180, 300), min_size=8),
        "dc5_s": case(coder, anchor, "dc5_s", 32, 64, 12000, 2000, (512, 1024)),                     # C5 of a 512x1024 input, train cfg
        "docstring": dict(rois=torch.tensor([[0., 0., 1., 1.], [0., 0., 1., 1.], [0., 0., 1., 1.], [5., 5., 5., 5.]]),
                          deltas=torch.tensor([[0., 0., 0., 0.], [1., 1., 1., 1.], [0., 0., 2., -1.], [0.7, -1.9, -0.5, 0.3]]),
                          max_shape=(32, 32, 3),
                          expected=torch.tensor([[0.0000, 0.0000, 1.0000, 1.0000], [0.1409, 0.1409, 2.8591, 2.8591],
                                                 [0.0000, 0.3161, 4.1945, 0.6839], [5.0000, 5.0000, 5.0000, 5.0000]])),
    }
    out["docstring"]["reference_output"] = coder.delta2bbox(out["docstring"]["rois"], out["docstring"]["deltas"], max_shape=(32, 32, 3))
    path = os.path.join(ROOT, "tests", "golden", "rpn_proposals.pt")
    torch.save(out, path)
    print({k: (tuple(v["dets"].shape) if "dets" in v else None) for k, v in out.items()}, os.path.getsize(path) / 1e3, "KB")


if __name__ == "__main__":
    main()
