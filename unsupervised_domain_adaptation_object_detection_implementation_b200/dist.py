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

    def __call__(self):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                self.flat[off:off + n].zero_()
            else:
                self.flat[off:off + n].copy_(p.grad.reshape(-1))
            off += n
        dist.all_reduce(self.flat)
        self.flat.div_(dist.get_world_size())
        off = 0
        for p in self.params:
            n = p.numel()
            g = self.flat[off:off + n].view(p.shape)
            if p.grad is None:
                p.grad = g.clone().to(p.dtype)
            else:
                p.grad.copy_(g.view_as(p.grad))
            off += n


class OverlappedGradAllReduce:
    """Gradient mean over ranks, overlapped with the rest of backward.

    Large parameters (the 411 MB FC1 gradient dominates this path) are all-reduced on a side stream the
    moment autograd has accumulated them (post-accumulate-grad hook), while RoIAlign backward and the
    image-head backward still run; everything small goes through one flat buffer at the end.  NCCL's
    AVG reduction does the division.  Works under CUDA-graph capture (the side stream forks and joins
    inside the capture)."""

    def __init__(self, params, big_numel=1 << 20):
        self.params = list(params)
        self.active = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.big = [p for p in self.params if p.numel() >= big_numel]
        self.small = [p for p in self.params if p.numel() < big_numel]
        self.flat = FlatGradAllReduce(self.small) if self.small else None
        self.comm_stream = None
        self._handles = []
        if self.active and self.params and self.params[0].is_cuda:
            # high priority: NCCL's CTAs are placed as soon as SMs free up instead of queueing behind the pending CTAs of
            # the backward kernels they are meant to overlap
            self.comm_stream = torch.cuda.Stream(device=self.params[0].device, priority=-1)
            for p in self.big:
                self._handles.append(p.register_post_accumulate_grad_hook(self._hook))
        self.op = dist.ReduceOp.AVG if (self.params and self.params[0].is_cuda) else dist.ReduceOp.SUM

    def _hook(self, p):
        cur = torch.cuda.current_stream()
        self.comm_stream.wait_stream(cur)
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


def max_over_ranks(value, device):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
