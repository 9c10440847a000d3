"""Fused SGD for the DA path.  Same update rule and constructor arguments as the torch.optim.SGD the
reference builds (mmdet/apis/train.py:127 from da_configs/faster_rcnn/faster_rcnn_r50_daf_c2f.py:8:
lr=1e-3, momentum=0.9, weight_decay=5e-4), executed by libda_b200's da_sgd_step: one pass over
(weight, grad, momentum buffer) that also refreshes the bf16 shadow weights of the tcgen05 engine."""
import torch

from . import functional as F_
from ._lib import lib, check


class FusedSGD:
    def __init__(self, params, lr=1e-3, momentum=0.9, weight_decay=0.0, shadow_bf16=True):
        self.params = [p for p in params if p.requires_grad]
        self.lr, self.momentum, self.weight_decay = float(lr), float(momentum), float(weight_decay)
        self.shadow_bf16 = shadow_bf16
        self.state = {}
        self.steps = 0

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    @torch.no_grad()
    def step(self):
        for p in self.params:
            g = p.grad
            if g is None:
                continue
            if not p.is_cuda or p.dtype != torch.float32:
                raise RuntimeError("FusedSGD updates fp32 CUDA parameters (no CPU fallback)")
            if not F_._dense_memory(p):
                raise RuntimeError("FusedSGD needs densely stored parameters")
            if g.dtype != torch.float32 or g.stride() != p.stride():
                g = torch.empty_like(p).copy_(g)      # same memory order as the parameter
            st = self.state.get(id(p))
            first = st is None
            if first:
                st = self.state[id(p)] = torch.empty_like(p)
            shadow = None
            if self.shadow_bf16 and p.dim() >= 2:
                shadow = F_.bf16_shadow(p)
            check(lib.da_sgd_step(F_._ptr(p), F_._ptr(g), F_._ptr(st), p.numel(), self.lr, self.momentum,
                                  self.weight_decay, int(first), F_._ptr(shadow), F_._stream()), "sgd_step")
        self.steps += 1
