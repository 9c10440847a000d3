"""Data-parallel plumbing of the DA path: one process per GPU, complete (source, target) pairs
per rank (never split a pair: SURVEY.md Q14), one flat-buffer gradient all-reduce per step over
NCCL/NVLink (gloo on CPU for tests).  The reference only wires MMDistributedDataParallel
(mmdet/apis/train.py:113-121); parameters that never receive gradients (Q9/Q10) are excluded
statically instead of relying on find_unused_parameters."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """torchrun / torch.distributed.run environment (RANK, LOCAL_RANK, WORLD_SIZE, MASTER_*)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


def shard_pairs(num_pairs, rank, world):
    """Contiguous, balanced assignment of whole pairs to ranks: returns the pair indices of `rank`."""
    base, rem = divmod(num_pairs, world)
    start = rank * base + min(rank, rem)
    return list(range(start, start + base + (1 if rank < rem else 0)))


def trainable_parameters(module, unused=()):
    skip = {id(p) for p in unused}
    return [p for p in module.parameters() if p.requires_grad and id(p) not in skip]


class FlatGradAllReduce:
    """Gradient mean over ranks through ONE contiguous buffer (a single NCCL all-reduce per step;
    NVSwitch makes message count, not link count, the cost)."""

    def __init__(self, params, dtype=torch.float32):
        self.params = list(params)
        self.numel = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(self.numel, dtype=dtype, device=dev)
        self._views = None
        self._fast_agreed = None
        self._live = None

    def __call__(self):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        fast = self.flat.is_cuda and all(p.grad is not None and p.grad.dtype == self.flat.dtype and p.grad.is_contiguous()
                                         for p in self.params)
        if self._fast_agreed is None:
            # the two paths issue different collectives: ALL ranks must take the same one.  Agreed once (first step), then
            # the choice is fixed; a rank that later falls off the fast path raises instead of deadlocking the others.
            t = torch.tensor([1.0 if fast else 0.0], device=self.flat.device)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            self._fast_agreed = bool(t.item() > 0)
        elif self._fast_agreed and not fast:
            raise RuntimeError("FlatGradAllReduce: a parameter lost its gradient on this rank after the ranks agreed on the "
                               "all-gradients path; list it in unused_parameters() or rebuild the reducer")
        if self._fast_agreed:
            # three launches in total: multi-tensor gather, NCCL AVG, multi-tensor scatter
            if self._views is None:
                self._views, off = [], 0
                for p in self.params:
                    self._views.append(self.flat[off:off + p.numel()])
                    off += p.numel()
            grads = [p.grad.reshape(-1) if p.grad.is_contiguous() else None for p in self.params]
            if all(g is not None for g in grads):
                torch._foreach_copy_(self._views, grads)
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)
                torch._foreach_copy_(grads, self._views)
                return
        # Slow path (a parameter without gradient on THIS rank, a dtype mismatch, or CPU/gloo).  A parameter whose gradient
        # is None on EVERY rank is skipped, exactly like on one GPU (torch.optim.SGD and FusedSGD skip grad=None, so no
        # weight decay / momentum is applied to it); a parameter that has a gradient on SOME rank gets zeros on the others.
        # Every rank runs the same collectives in the same order with the same op.
        world = dist.get_world_size()
        if self._live is None:      # agreed once, outside any graph capture (first eager step); static afterwards
            has = torch.tensor([0.0 if p.grad is None else 1.0 for p in self.params], dtype=torch.float32, device=self.flat.device)
            if len(self.params):
                dist.all_reduce(has, op=dist.ReduceOp.SUM)
            self._live = (has > 0).tolist()
        anywhere = self._live
        for p, live in zip(self.params, anywhere):
            if p.grad is not None and not live:
                raise RuntimeError("FlatGradAllReduce: a parameter that had no gradient on any rank at the first step has one now; "
                                   "rebuild the reducer")
        off = 0
        for p, live in zip(self.params, anywhere):
            n = p.numel()
            if p.grad is None or not live:
                self.flat[off:off + n].zero_()
            else:
                self.flat[off:off + n].copy_(p.grad.reshape(-1))
            off += n
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        self.flat.div_(world)
        off = 0
        for p, live in zip(self.params, anywhere):
            n = p.numel()
            if live:
                g = self.flat[off:off + n].view(p.shape)
                if p.grad is None:
                    p.grad = g.clone().to(p.dtype)
                else:
                    p.grad.copy_(g.view_as(p.grad))
            off += n


class OverlappedGradAllReduce:
    """Gradient mean over ranks, overlapped with the rest of backward.

    Large parameters (the 411 MB FC1 gradient dominates this path: 0.84 ms over NVLink at N=2) are all-reduced on a
    side stream, everything small goes through one flat buffer at the end.  NCCL's AVG reduction does the division.
    Two things make the overlap effective:
      * the all-reduce of a layer starts right after its WEIGHT-gradient kernel (functional.WGRAD_HOOK records an
        event there), so the layer's own data gradient already runs under it;
      * NCCL keeps `nccl_sms` SMs busy for the whole transfer; a persistent GEMM grid of one CTA per physical SM would
        then run in two waves, so the SM budget of the persistent kernels is lowered (da_set_sm_limit) from the moment
        the first all-reduce is enqueued until the join.
    Works under CUDA-graph capture (the side stream forks and joins inside the capture; grid sizes are baked in)."""

    def __init__(self, params, big_numel=1 << 20, early_numel=1 << 25, nccl_sms=32):
        self.params = list(params)
        self.active = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.big = [p for p in self.params if p.numel() >= big_numel]
        self.small = [p for p in self.params if p.numel() < big_numel]
        self.flat = FlatGradAllReduce(self.small) if self.small else None
        self.early_numel, self.nccl_sms = early_numel, nccl_sms
        self.comm_stream = None
        self._handles = []
        self._events = {}
        self._limited = False
        self._F = None
        if self.active and self.params and self.params[0].is_cuda:
            # high priority: NCCL's CTAs are placed as soon as SMs free up instead of queueing behind the pending CTAs of
            # the backward kernels they are meant to overlap
            self.comm_stream = torch.cuda.Stream(device=self.params[0].device, priority=-1)
            from . import functional as F_
            self._F = F_
            F_.WGRAD_HOOK = self._after_wgrad
            self._sms = torch.cuda.get_device_properties(self.params[0].device).multi_processor_count
            for p in self.big:
                self._handles.append(p.register_post_accumulate_grad_hook(self._hook))
        self.op = dist.ReduceOp.AVG if (self.params and self.params[0].is_cuda) else dist.ReduceOp.SUM

    def _after_wgrad(self, dw):
        """Called by the layer right after its weight-gradient kernel: event for the early all-reduce, and from here on
        the persistent kernels leave room for NCCL."""
        if dw.numel() < self.early_numel:
            return
        ev = torch.cuda.Event()
        ev.record()
        self._events[dw.data_ptr()] = ev
        if not self._limited and self.nccl_sms > 0:
            self._F.set_sm_limit(self._sms - self.nccl_sms)
            self._limited = True

    def _hook(self, p):
        ev = self._events.pop(p.grad.data_ptr(), None)     # only valid if autograd kept the layer's tensor as .grad
        if ev is not None:
            self.comm_stream.wait_event(ev)
        else:
            self.comm_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm_stream):
            dist.all_reduce(p.grad, op=self.op)

    def __call__(self):
        """Call after backward: reduces the small parameters and joins the side stream."""
        if not self.active:
            return
        if self.comm_stream is None:          # CPU / gloo: no overlap, no AVG
            FlatGradAllReduce(self.params)()
            return
        if self.flat is not None:
            self.flat()
        torch.cuda.current_stream().wait_stream(self.comm_stream)
        self._events.clear()
        if self._limited:
            self._F.set_sm_limit(0)
            self._limited = False


def max_over_ranks(value, device):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
