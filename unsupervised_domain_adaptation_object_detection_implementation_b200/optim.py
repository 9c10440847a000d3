"""Fused SGD for the DA path.  Same update rule and constructor arguments as the torch.optim.SGD the
reference builds (mmdet/apis/train.py:127 from da_configs/faster_rcnn/faster_rcnn_r50_daf_c2f.py:8:
lr=1e-3, momentum=0.9, weight_decay=5e-4), executed by libda_b200's da_sgd_step: one pass over
(weight, grad, momentum buffer) that also refreshes the bf16 shadow weights of the tcgen05 engine."""
import torch

from . import functional as F_
from ._lib import lib, check, SgdFuse as _lib_SgdFuse

_CHUNK = 8192    # DA_SGD_CHUNK (include/da_b200.h)


def _same_layout(a, b):
    """Same memory order (strides of size-1 dimensions carry no information)."""
    return a.shape == b.shape and all(n == 1 or sa == sb for n, sa, sb in zip(a.shape, a.stride(), b.stride()))


class StepLrSchedule:
    """mmcv's StepLrUpdaterHook with linear warm-up, as the DA configs use it (da_configs/_base_/schedules/schedule_1x.py:
    lr_config = dict(policy='step', warmup='linear', warmup_iters=500, warmup_ratio=0.0001, step=[6, 8]); the DAF recipe
    overrides step=[9], faster_rcnn_r50_daf_c2f.py:13-18).  lr(iter, epoch) =
        base * gamma ** #{s in step : epoch >= s}                         regular
        regular * (1 - (1 - iter / warmup_iters) * (1 - warmup_ratio))    while iter < warmup_iters (linear warm-up)
    Hyper-parameters reach the kernels as launch arguments: assign `optimizer.lr` every iteration of an EAGER loop; a
    captured CUDA graph or the peer optimizer bakes the value in, so re-capture when the schedule changes it."""

    def __init__(self, base_lr, lr_config=None):
        cfg = dict(lr_config or {})
        if cfg.get("policy", "step") != "step":
            raise NotImplementedError("the DA configs use policy='step' (da_configs/_base_/schedules/schedule_1x.py)")
        self.base_lr = float(base_lr)
        step = cfg.get("step", [])
        self.steps = [int(step)] if isinstance(step, int) else [int(s) for s in step]
        self.gamma = float(cfg.get("gamma", 0.1))
        self.warmup = cfg.get("warmup")
        if self.warmup not in (None, "linear", "constant"):
            raise NotImplementedError(f"warmup={self.warmup!r} (mmcv supports constant/linear/exp; the DA configs use linear)")
        self.warmup_iters = int(cfg.get("warmup_iters", 0))
        self.warmup_ratio = float(cfg.get("warmup_ratio", 0.1))

    def __call__(self, it, epoch):
        """Learning rate of iteration `it` (0-based, global) in epoch `epoch` (0-based)."""
        lr = self.base_lr * self.gamma ** sum(1 for s in self.steps if epoch >= s)
        if self.warmup is not None and it < self.warmup_iters:
            if self.warmup == "constant":
                return lr * self.warmup_ratio
            return lr * (1.0 - (1.0 - it / self.warmup_iters) * (1.0 - self.warmup_ratio))
        return lr


class FusedSGD:
    def __init__(self, params, lr=1e-3, momentum=0.9, weight_decay=0.0, shadow_bf16=True, fuse_wgrad=()):
        """fuse_wgrad: parameters (weights of functional.dense_layer layers on the bf16 tensor-core engine) whose update
        is applied by the epilogue of their weight-gradient kernel during backward (da_conv_backward_weight_sgd): their
        gradient is never materialised and `step()` skips them.  One backward per step; one GPU (no gradient exchange)."""
        self.params = [p for p in params if p.requires_grad]
        self._order = list(self.params)           # constructor order = the index space of state_dict() (torch.optim.SGD layout)
        self.fused = []
        for p in fuse_wgrad:
            self._register_fused(p)
        self.lr, self.momentum, self.weight_decay = float(lr), float(momentum), float(weight_decay)
        self.shadow_bf16 = shadow_bf16
        self.state = {}
        self.steps = 0
        self._chunks, self._chunk_key, self._keep, self._host, self._table, self._copied = None, None, None, None, None, None
        self._spare, self._captured = [], []

    def _register_fused(self, p):
        if not (p.is_cuda and p.dtype == torch.float32 and F_._dense_memory(p) and p.dim() >= 2):
            raise RuntimeError("FusedSGD(fuse_wgrad): densely stored fp32 CUDA weights only")
        self.params = [q for q in self.params if q is not p]
        buf = torch.zeros_like(p)
        shadow = F_.bf16_shadow(p)
        state = {"calls": 0}
        opt = self

        class _Fused(F_.ManagedWeight):
            def fuse(self_inner):
                rec = _lib_SgdFuse(p.data_ptr(), buf.data_ptr(), shadow.data_ptr(), opt.lr, opt.momentum, opt.weight_decay,
                                   int(state["calls"] == 0))
                state["calls"] += 1
                return rec

        F_.MANAGED_WGRAD[id(p)] = _Fused()
        self.fused.append((p, buf, shadow))
        self._fused_state = getattr(self, "_fused_state", {})
        self._fused_state[id(p)] = state
        if all(q is not p for q in self._order):
            self._order.append(p)

    # ---- checkpointing: the layout of torch.optim.SGD.state_dict() (what mmcv's save_checkpoint stores for the reference,
    # mmdet/apis/train.py:127 + mmcv/runner/checkpoint.py), so that optimizer state moves both ways
    def state_dict(self):
        state = {}
        fused = {id(p): (buf, self._fused_state[id(p)]) for p, buf, _ in self.fused}
        for i, p in enumerate(self._order):
            if id(p) in fused:
                buf, st = fused[id(p)]
                if st["calls"] > 0:
                    state[i] = {"momentum_buffer": buf.detach().clone()}
            elif id(p) in self.state:
                state[i] = {"momentum_buffer": self.state[id(p)].detach().clone()}
        group = {"lr": self.lr, "momentum": self.momentum, "dampening": 0, "weight_decay": self.weight_decay, "nesterov": False,
                 "maximize": False, "foreach": None, "differentiable": False, "params": list(range(len(self._order)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        groups = sd.get("param_groups", [])
        if groups:
            g = groups[0]
            self.lr, self.momentum = float(g.get("lr", self.lr)), float(g.get("momentum", self.momentum))
            self.weight_decay = float(g.get("weight_decay", self.weight_decay))
        fused = {id(p): (buf, self._fused_state[id(p)]) for p, buf, _ in self.fused}
        for i, st in sd.get("state", {}).items():
            i = int(i)
            buf = st.get("momentum_buffer") if isinstance(st, dict) else None
            if buf is None or i >= len(self._order):
                continue                                  # a parameter that never had a gradient has no momentum yet
            p = self._order[i]
            if buf.numel() != p.numel():
                raise RuntimeError(f"FusedSGD.load_state_dict: momentum of parameter {i} has {buf.numel()} elements, expected {p.numel()}")
            if id(p) in fused:
                fused[id(p)][0].copy_(buf.to(p.device).view_as(p))
                fused[id(p)][1]["calls"] = max(1, fused[id(p)][1]["calls"])
            else:
                tgt = self.state.get(id(p))
                if tgt is None:
                    tgt = self.state[id(p)] = torch.empty_like(p)
                tgt.copy_(buf.to(p.device).reshape(p.shape))

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    @torch.no_grad()
    def step(self):
        """One launch for all parameters (da_sgd_step_multi): the per-tensor records {w, grad, buf, shadow, n, first}
        go to the device as one small table; the (tensor, chunk) tiling is built once."""
        rows, keep = [], []
        for p in self.params:
            g = p.grad
            if g is None:
                continue
            if not p.is_cuda or p.dtype != torch.float32:
                raise RuntimeError("FusedSGD updates fp32 CUDA parameters (no CPU fallback)")
            if not F_._dense_memory(p):
                raise RuntimeError("FusedSGD needs densely stored parameters")
            if g.dtype != torch.float32 or not _same_layout(g, p):
                g = torch.empty_like(p).copy_(g)      # same memory order as the parameter
                keep.append(g)
            st = self.state.get(id(p))
            first = st is None
            if first:
                st = self.state[id(p)] = torch.empty_like(p)
            shadow = F_.bf16_shadow(p) if (self.shadow_bf16 and p.dim() >= 2) else None
            for t in (p, g, st):
                if t.data_ptr() % 16:
                    raise RuntimeError("FusedSGD: tensors must be 16-byte aligned")
            rows.append((p.data_ptr(), g.data_ptr(), st.data_ptr(), shadow.data_ptr() if shadow is not None else 0,
                         p.numel(), int(first)))
        if not rows:
            self.steps += 1
            return
        dev = self.params[0].device
        sizes = tuple(r[4] for r in rows)
        if self._chunks is None or self._chunk_key != sizes:
            ch = [(i, c) for i, n in enumerate(sizes) for c in range((n + _CHUNK - 1) // _CHUNK)]
            self._chunks = torch.tensor(ch, dtype=torch.int32).reshape(-1, 2).to(dev)
            self._chunk_key = sizes
        # table of da_sgd_entry records (6 x int64 each; first_step sits in the low 32 bits of the last word).  The pinned
        # staging buffer and the device table are allocated once (first eager step), so a captured step only holds a copy node.
        if self._host is None or self._host[0].shape[0] < len(rows):
            # 2 rotating staging buffers for eager steps + spares that a CUDA-graph capture takes for good (a captured
            # copy node re-reads its host buffer at every replay, so that buffer must never be rewritten)
            self._host = [torch.empty((len(self.params), 6), dtype=torch.int64).pin_memory() for _ in range(2)]
            self._spare = [torch.empty((len(self.params), 6), dtype=torch.int64).pin_memory() for _ in range(4)]
            self._table = torch.empty((len(self.params), 6), dtype=torch.int64, device=dev)
            self._copied = [None, None]
        capturing = torch.cuda.is_current_stream_capturing()
        if capturing:
            if not self._spare:
                raise RuntimeError("FusedSGD: more than 4 graph captures of step(); run one eager step first / raise the spare count")
            stage = self._spare.pop()
            self._captured.append(stage)
        else:
            slot = self.steps & 1
            if self._copied[slot] is not None:
                self._copied[slot].synchronize()          # the copy that last read this staging buffer has run
            stage = self._host[slot]
        stage[:len(rows)] = torch.tensor(rows, dtype=torch.int64)
        table = self._table
        table[:len(rows)].copy_(stage[:len(rows)], non_blocking=True)
        if not capturing:
            self._copied[slot] = torch.cuda.Event()
            self._copied[slot].record()
        self._keep = keep   # alive until the next step
        check(lib.da_sgd_step_multi(F_._ptr(table), len(rows), F_._ptr(self._chunks), self._chunks.shape[0], self.lr,
                                    self.momentum, self.weight_decay, F_._stream()), "sgd_step_multi")
        self.steps += 1
