"""RPN proposal stage per image at the train configuration (C5 of a 1024x2048 input: 64x128 cells x 15 anchors, nms_pre 12000,
max_per_img 2000, IoU 0.7): the native stage (functional.rpn_proposals) next to the torch-op path the detector shell used before
(sigmoid / topk / gathers / decode in ATen + torchvision.ops.nms).  CUDA events, median of 20."""
import json, sys, torch
sys.path.insert(0, ".")
from unsupervised_domain_adaptation_object_detection_implementation_b200 import functional as F_, detection
from oracle import rpn_oracle
dev = "cuda"
A, H, W = 15, 64, 128
g = torch.Generator(device=dev).manual_seed(0)
cls = torch.randn(A, H, W, device=dev, generator=g) * 2
reg = torch.randn(4 * A, H, W, device=dev, generator=g) * 0.5
base = rpn_oracle.base_anchors(16, [0.5, 1.0, 2.0], [2, 4, 8, 16, 32])
head = detection.RPNHeadDA(64, feat_channels=64, anchor_generator=dict(scales=[2, 4, 8, 16, 32], ratios=[0.5, 1.0, 2.0], strides=[16]))
anchors = head.anchor_generator.grid_anchors((H, W), 0, dev)
cfg = dict(nms_pre=12000, max_per_img=2000, nms=dict(iou_threshold=0.7), min_bbox_size=0)
def torch_path():
    scores = cls.permute(1, 2, 0).reshape(-1).sigmoid()
    deltas = reg.permute(1, 2, 0).reshape(-1, 4)
    s, idx = scores.topk(12000)
    boxes = head.bbox_coder.decode(anchors[idx], deltas[idx], max_shape=(1024, 2048))
    keep = ((boxes[:, 2] - boxes[:, 0]) > 0) & ((boxes[:, 3] - boxes[:, 1]) > 0)
    boxes, s = boxes[keep], s[keep]
    from torchvision.ops import nms
    k = nms(boxes, s, 0.7)[:2000]
    return torch.cat([boxes[k], s[k, None]], -1)
def native():
    return F_.rpn_proposals(cls, reg, base, 16, (1024, 2048), 12000, 2000, 0.7, 0.0)
def ev(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[n // 2]
F_.lib.da_launch_count_reset()
native(); torch.cuda.synchronize()
launches = int(F_.lib.da_launch_count())
t_native, t_torch = ev(native), ev(torch_path)
n_native = int(native()[1]); n_torch = torch_path().shape[0]
print(json.dumps({"what": "RPN proposal stage, one image, 122880 candidates -> 12000 ranked -> NMS 0.7 -> 2000", "native_ms": round(t_native, 3),
                  "torch_ops_plus_torchvision_nms_ms": round(t_torch, 3), "our_launches": launches, "proposals_native": n_native,
                  "proposals_torch_path": n_torch,
                  "note": "native: 4 of our kernels + torch.sort; the torch path includes the boolean-mask compaction and NMS host syncs"}))
