"""Peer-memory optimizer for the large tensors of the DA path (N GPUs of one NVSwitch box).

The reference trains data-parallel through MMDistributedDataParallel + torch.optim.SGD
(mmdet/apis/train.py:113-127): all-reduce every gradient, then every rank repeats the whole update.  For
FC1 of the shared bbox head (103 M parameters) that is a 411 MB all-reduce and a 2.26 GB optimizer pass
per step and rank.  `PeerShardedSGD` replaces both for the tensors it manages (`da_sgd_step_peer`,
csrc/peer_sgd.cu): rank r owns slice r of the tensor; the copy engines push slice r of every rank's gradient
into rank r's staging area over NVLink while the backward pass continues, one all-local kernel averages the
staged slices in rank order and applies the SGD rule, and the copy engines push the refreshed bf16 operand
slice to every rank (optionally under the beginning of the next step: `deferred_publish`).  Momentum is
stored sharded; the fp32 master of a rank is current on its own slice only unless `share_master=True`
(`gather_master()` assembles it for checkpoints).

Buffers that peers touch (gradient, bf16 operand copy, flag words) are cudaMalloc'ed blocks exchanged as
CUDA IPC handles through torch.distributed; nothing here goes through NCCL.
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib, functional as F_
from ._lib import check, lib


class PeerBlock:
    """One cudaMalloc'ed, zero-filled device block (da_peer_alloc) that torch can view and peers can map."""

    def __init__(self, nbytes, device):
        self.nbytes, self.device = int(nbytes), torch.device(device)
        out = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(lib.da_peer_alloc(self.nbytes, ctypes.byref(out)), "peer_alloc")
        self.ptr = out.value

    def handle(self):
        buf = (ctypes.c_ubyte * _lib.DA_PEER_HANDLE_BYTES)()
        check(lib.da_peer_export(self.ptr, buf), "peer_export")
        return bytes(buf)

    def tensor(self, dtype, numel):
        """Flat torch view of the block (bf16 goes through int16: the CUDA array interface has no bf16)."""
        carrier = torch.int16 if dtype == torch.bfloat16 else dtype
        typestr = {torch.float32: "<f4", torch.int32: "<i4", torch.int16: "<i2"}[carrier]
        owner = self

        class _View:
            __cuda_array_interface__ = {"shape": (int(numel),), "typestr": typestr, "data": (owner.ptr, False), "version": 2}
            _keep = owner

        t = torch.as_tensor(_View(), device=self.device)
        return t.view(torch.bfloat16) if dtype == torch.bfloat16 else t


def open_peer(handle_bytes):
    buf = (ctypes.c_ubyte * _lib.DA_PEER_HANDLE_BYTES).from_buffer_copy(handle_bytes)
    out = ctypes.c_void_p()
    check(lib.da_peer_open(buf, ctypes.byref(out)), "peer_open")
    return out.value


def slice_bounds(n, world, rank):
    """The slice of rank `rank` (same rule as the kernel: multiples of 1024 elements, remainder to the last ranks)."""
    per = ((n + world - 1) // world + 1023) // 1024 * 1024
    lo = min(per * rank, n)
    return lo, min(lo + per, n), per


def make_args(w, momentum_shard, grads, shadows, masters, flags, local_state, n, world, rank):
    """da_peer_sgd_args from raw device pointers (ints; 0/None = NULL)."""
    a = _lib.PeerSgdArgs()
    a.w, a.momentum_shard, a.local_state = w, momentum_shard, local_state
    for r in range(world):
        a.grad[r] = grads[r]
        a.w_bf16[r] = shadows[r] or None
        a.w_f32[r] = (masters[r] if masters else 0) or None
        a.flags[r] = flags[r]
    a.n, a.world, a.rank = int(n), int(world), int(rank)
    return a


class _Managed:
    __slots__ = ("param", "n", "grad", "shadow", "flags", "state", "momentum", "args", "calls", "blocks", "pending", "ptrs",
                 "staging", "lo", "hi", "per")


class PeerShardedSGD:
    """SGD (momentum, weight decay; the reference recipe) for `params`, fused with the gradient mean over ranks and
    launched from INSIDE backward: the exchange starts when the layer's weight gradient exists and the update runs on a
    side stream as soon as the layer's own backward kernels are enqueued, under the rest of the backward pass.
    (Without a process group it degenerates to a one-rank update on the side stream; on one GPU that overlap was measured
    and does not pay - optim.FusedSGD(fuse_wgrad=...) is the one-GPU form.)

    Usage per step:  forward/backward (the managed layers call back from their backward: `_after_wgrad` when the
    weight-gradient kernel is enqueued, `_layer_done` when the layer's last backward kernel is) -> `join()` before
    the next forward.  One backward per step: the weight-gradient kernel overwrites the gradient buffer.

    transport="copy" (default): the copy engines push slice r of the local gradient into rank r's staging area as soon
    as the weight gradient exists, the update kernel touches local memory only, the copy engines push the refreshed
    bf16 slice to every rank.  transport="stores": one kernel does everything with SM-issued P2P loads and stores."""

    def __init__(self, params, lr=1e-3, momentum=0.9, weight_decay=0.0, max_ctas=0, reserve_sms=0, share_master=False,
                 transport="copy", deferred_publish=False, group=None):
        if transport not in ("copy", "stores"):
            raise ValueError("transport must be 'copy' or 'stores'")
        if dist.is_available() and dist.is_initialized():
            self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        else:       # one GPU: no peers, the update still runs on the side stream under the rest of backward
            self.world, self.rank = 1, 0
        if self.world > _lib.DA_MAX_PEERS:
            raise RuntimeError(f"PeerShardedSGD: at most {_lib.DA_MAX_PEERS} ranks (one NVSwitch box)")
        self.lr, self.momentum, self.weight_decay = float(lr), float(momentum), float(weight_decay)
        self.max_ctas, self.share_master, self.group = int(max_ctas), bool(share_master), group
        self.reserve_sms, self.transport = int(reserve_sms), transport
        # deferred publish (copy transport): the pushes of the refreshed bf16 slices are NOT part of the step; the caller
        # enqueues them with publish() right after the step (outside a captured graph) and they run under the beginning of
        # the next step, whose first use of the operand copy is preceded by the wait for every peer's pushes
        # (ManagedWeight.before_forward).  Per step: forward/backward -> join() -> publish().
        self.deferred = bool(deferred_publish) and transport == "copy"
        self.items = []
        params = [p for p in params if p.requires_grad]
        if not params:
            return
        dev = params[0].device
        self.device = dev
        self.comm_stream = torch.cuda.Stream(device=dev, priority=-1)
        self._sms = torch.cuda.get_device_properties(dev).multi_processor_count
        self._limited = False
        local_handles = []
        copy = transport == "copy"
        for p in params:
            if p.dtype != torch.float32 or not p.is_cuda or not F_._dense_memory(p):
                raise RuntimeError("PeerShardedSGD manages densely stored fp32 CUDA parameters")
            if self.world > 1:
                dist.broadcast(p.data, 0, group=group)      # identical masters to start from
            m = _Managed()
            m.param, m.n, m.calls, m.pending = p, p.numel(), 0, False
            m.lo, m.hi, m.per = slice_bounds(m.n, self.world, self.rank)
            # blocks peers map: [0] gradient (stores) or staging area world x per (copy), [1] bf16 operand copy, [2] flags,
            # [3] fp32 master (share_master)
            blocks = [PeerBlock(4 * (self.world * m.per if copy else m.n), dev), PeerBlock(2 * m.n, dev),
                      PeerBlock(4 * _lib.DA_PEER_FLAG_INTS, dev)]
            if self.share_master:
                blocks.append(PeerBlock(4 * m.n, dev))
            m.blocks = blocks
            if copy:
                m.staging = blocks[0]
                m.grad = torch.zeros(m.n, dtype=torch.float32, device=dev)     # local only: peers never read it
            else:
                m.staging = None
                m.grad = blocks[0].tensor(torch.float32, m.n)
            m.shadow = blocks[1].tensor(torch.bfloat16, m.n)
            m.flags = blocks[2].tensor(torch.int32, _lib.DA_PEER_FLAG_INTS)
            if self.share_master:                            # re-home the master so that peers can write into it
                master = blocks[3].tensor(torch.float32, m.n).as_strided(p.shape, p.stride())
                master.copy_(p.data)
                p.data = master
            m.state = torch.zeros(4, dtype=torch.int32, device=dev)
            m.momentum = torch.zeros(max(m.per, 8), dtype=torch.float32, device=dev)
            # the operand copy every layer call will use from now on (functional.bf16_shadow's cache)
            check(lib.da_cast(F_._ptr(p.data), _lib.DA_F32, F_._ptr(m.shadow), _lib.DA_BF16, m.n, F_._stream()), "cast")
            p._da_shadow = (p._version, m.shadow.as_strided(p.shape, p.stride()))
            local_handles.append([b.handle() for b in blocks])
            self.items.append(m)
        torch.cuda.synchronize(dev)
        gathered = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(gathered, local_handles, group=group)
        for k, m in enumerate(self.items):
            ptrs = [[b.ptr for b in m.blocks] if r == self.rank else [open_peer(h) for h in gathered[r][k]]
                    for r in range(self.world)]
            m.ptrs = ptrs
            flags = [q[2] for q in ptrs]
            if copy:
                # grad[r] + i must address element i (absolute index, lo <= i < hi) of what rank r pushed: slot r of the
                # LOCAL staging area; the own gradient is read in place
                base = m.blocks[0].ptr
                grads = [m.grad.data_ptr() if r == self.rank else base + 4 * (r * m.per - m.lo) for r in range(self.world)]
                shadows = [m.shadow.data_ptr() if r == self.rank else 0 for r in range(self.world)]
                m.args = make_args(m.param.data_ptr(), m.momentum.data_ptr(), grads, shadows, None, flags,
                                   m.state.data_ptr(), m.n, self.world, self.rank)
            else:
                m.args = make_args(m.param.data_ptr(), m.momentum.data_ptr(), [q[0] for q in ptrs], [q[1] for q in ptrs],
                                   [q[3] for q in ptrs] if self.share_master else None, flags,
                                   m.state.data_ptr(), m.n, self.world, self.rank)
            mw = F_.ManagedWeight()
            mw.grad = m.grad
            mw.after_wgrad = (lambda m=m: self._after_wgrad(m))
            mw.layer_done = (lambda m=m: self._layer_done(m))
            if self.deferred and self.world > 1:
                mw.before_forward = (lambda m=m: self._before_forward(m))
            F_.MANAGED_WGRAD[id(m.param)] = mw
        if self.world > 1:
            dist.barrier(group=group)                         # every rank has mapped every block before the first step

    def _comm_after_current(self):
        ev = torch.cuda.Event()
        ev.record()
        self.comm_stream.wait_event(ev)
        return ctypes.c_void_p(self.comm_stream.cuda_stream)

    def _after_wgrad(self, m):
        """The weight-gradient kernel is enqueued: start pushing the slices the other ranks own (copy engines)."""
        if m.pending:
            raise RuntimeError("PeerShardedSGD: two backward passes through a managed layer in one step "
                               "(gradient accumulation is not supported on the peer path)")
        m.pending = True
        if self.transport != "copy":
            return
        st = self._comm_after_current()
        for k in range(1, self.world):
            r = (self.rank + k) % self.world
            lo, hi, _ = slice_bounds(m.n, self.world, r)
            check(lib.da_peer_copy(m.ptrs[r][0] + 4 * self.rank * m.per, m.grad.data_ptr() + 4 * lo, 4 * (hi - lo), st), "peer_copy")

    def _layer_done(self, m):
        """The layer's backward is enqueued (it no longer reads its operand copy): update the own slice and publish it."""
        st = self._comm_after_current()
        first = int(m.calls == 0)
        if self.transport == "copy":
            check(lib.da_sgd_step_peer(ctypes.byref(m.args), self.lr, self.momentum, self.weight_decay, first, self.max_ctas,
                                       _lib.DA_PEER_PUBLISH_BY_CALLER, st), "sgd_step_peer")
            if not (self.deferred and self.world > 1):      # deferred: publish() enqueues the pushes after the step
                self._push_slices(m, st)
                if self.world > 1:
                    check(lib.da_peer_publish_done(ctypes.byref(m.args), st), "peer_publish_done")
        else:
            check(lib.da_sgd_step_peer(ctypes.byref(m.args), self.lr, self.momentum, self.weight_decay, first, self.max_ctas,
                                       _lib.DA_PEER_PUBLISH_STORES, st), "sgd_step_peer")
        m.calls += 1
        if not self._limited and self.reserve_sms > 0:
            F_.set_sm_limit(self._sms - self.reserve_sms)    # persistent kernels leave room for the peer kernel
            self._limited = True

    def _push_slices(self, m, st):
        for k in range(1, self.world):
            r = (self.rank + k) % self.world
            check(lib.da_peer_copy(m.ptrs[r][1] + 2 * m.lo, m.shadow.data_ptr() + 2 * m.lo, 2 * (m.hi - m.lo), st), "peer_copy")
            if self.share_master:
                check(lib.da_peer_copy(m.ptrs[r][3] + 4 * m.lo, m.param.data_ptr() + 4 * m.lo, 4 * (m.hi - m.lo), st), "peer_copy")

    def _before_forward(self, m):
        """In front of the layer's forward: every peer's pushes of the previous step have landed in this rank's copy."""
        check(lib.da_peer_wait_done(ctypes.byref(m.args), F_._stream()), "peer_wait_done")

    def publish(self):
        """Deferred publish: enqueue the pushes of this step's refreshed slices + the done signal on the side stream, ordered
        after everything enqueued on the current stream so far (call it right after the step / after replaying its graph;
        not capturable on purpose: it must not be joined by the step)."""
        if not (self.deferred and self.world > 1 and self.items):
            return
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("PeerShardedSGD.publish() belongs outside the captured step")
        st = self._comm_after_current()
        for m in self.items:
            self._push_slices(m, st)
            check(lib.da_peer_signal_done(ctypes.byref(m.args), st), "peer_signal_done")

    def join(self):
        """The step's peer kernels are ordered before whatever the current stream does next."""
        if not self.items:
            return
        torch.cuda.current_stream().wait_stream(self.comm_stream)
        for m in self.items:
            m.pending = False
        if self._limited:
            F_.set_sm_limit(0)
            self._limited = False

    def check_errors(self):
        """Raises if a cross-GPU barrier timed out (local_state[2])."""
        for m in self.items:
            err = int(m.state[2].item())
            if err:
                what = {1: "barrier-in", 2: "barrier-out", 3: "wait for the previous publish"}.get(err, str(err))
                raise RuntimeError(f"PeerShardedSGD: {what} timed out on rank {self.rank}")

    @torch.no_grad()
    def gather_master(self):
        """Assemble the full fp32 master on every rank (checkpointing); a no-op with share_master."""
        if self.share_master or self.world == 1:
            return
        for m in self.items:
            flat = m.param.data.as_strided((m.n,), (1,))
            lo, hi, per = slice_bounds(m.n, self.world, self.rank)
            pieces = [torch.empty(per, dtype=torch.float32, device=self.device) for _ in range(self.world)]
            mine = torch.zeros(per, dtype=torch.float32, device=self.device)
            mine[:hi - lo] = flat[lo:hi]
            dist.all_gather(pieces, mine, group=self.group)
            for r, piece in enumerate(pieces):
                a, b, _ = slice_bounds(m.n, self.world, r)
                flat[a:b] = piece[:b - a]

    def managed_parameters(self):
        return [m.param for m in self.items]
