"""GPU tests of the peer-memory optimizer kernel (da_sgd_step_peer, csrc/peer_sgd.cu).

The whole cross-rank protocol (flag barriers, slice ownership, gradient mean in rank order, SGD rule, operand
broadcast) is exercised on ONE device: `world` virtual ranks live in one process, each with its own buffers and its
own stream, and point at each other's buffers directly (CUDA IPC is only the transport of those pointers between
processes; tools/peer_check.py covers it under torchrun on >= 2 GPUs and is run by the last test when they exist).
Expected values: mean of the gradients in rank order, then da_sgd_step (already pinned to torch.optim.SGD by
test_gpu_parity.py) on the full tensor.  Everything is compared bit for bit.
"""
import ctypes
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

from unsupervised_domain_adaptation_object_detection_implementation_b200 import _lib, functional as F_, peer  # noqa: E402
from unsupervised_domain_adaptation_object_detection_implementation_b200._lib import check, lib  # noqa: E402

DEV = "cuda"
LR, MU, WD = 0.05, 0.9, 5e-4


def _expected(w0, grads_per_step, world):
    """Reference trajectory: (masters per step, bf16 copies per step) from da_sgd_step on the averaged gradient."""
    w = w0.clone()
    buf = torch.zeros_like(w)
    shadow = torch.empty_like(w, dtype=torch.bfloat16)
    inv = torch.tensor(1.0 / world, dtype=torch.float32, device=DEV)
    outs = []
    for s, grads in enumerate(grads_per_step):
        g = grads[0].clone()
        for r in range(1, world):
            g = g + grads[r]
        g = g * inv
        check(lib.da_sgd_step(F_._ptr(w), F_._ptr(g), F_._ptr(buf), w.numel(), LR, MU, WD, int(s == 0), F_._ptr(shadow), None), "sgd")
        torch.cuda.synchronize()
        outs.append((w.clone(), shadow.clone()))
    return outs


@pytest.mark.parametrize("world,n,share_master", [(1, 8192 + 5, False), (2, 3 * 1024 * 37 + 13, False), (3, 1024 * 50 + 3, True),
                                                  (4, 1024 * 1024 + 1024, False), (8, 8 * 1024 * 9, True)])
def test_peer_sgd_virtual_ranks_match_allreduce_plus_sgd(world, n, share_master):
    g = torch.Generator(device=DEV).manual_seed(world * 1000 + n % 97)
    w0 = torch.randn(n, device=DEV, generator=g)
    steps = 3
    grads_per_step = [[torch.randn(n, device=DEV, generator=g) for _ in range(world)] for _ in range(steps)]
    exp = _expected(w0, grads_per_step, world)

    masters = [w0.clone() for _ in range(world)]
    grads = [torch.empty(n, device=DEV) for _ in range(world)]
    shadows = [torch.zeros(n, device=DEV, dtype=torch.bfloat16) for _ in range(world)]
    flags = [torch.zeros(_lib.DA_PEER_FLAG_INTS, dtype=torch.int32, device=DEV) for _ in range(world)]
    states = [torch.zeros(4, dtype=torch.int32, device=DEV) for _ in range(world)]
    per = peer.slice_bounds(n, world, 0)[2]
    moms = [torch.zeros(max(per, 8), device=DEV) for _ in range(world)]
    args = [peer.make_args(masters[r].data_ptr(), moms[r].data_ptr(), [t.data_ptr() for t in grads],
                           [t.data_ptr() for t in shadows], [t.data_ptr() for t in masters] if share_master else None,
                           [t.data_ptr() for t in flags], states[r].data_ptr(), n, world, r) for r in range(world)]
    streams = [torch.cuda.Stream() for _ in range(world)]
    for s in range(steps):
        for r in range(world):
            grads[r].copy_(grads_per_step[s][r])
        torch.cuda.synchronize()
        for r in range(world):          # all virtual ranks must be co-resident: 8 x 8 CTAs at most
            check(lib.da_sgd_step_peer(ctypes.byref(args[r]), LR, MU, WD, int(s == 0), 8, _lib.DA_PEER_PUBLISH_STORES,
                                       ctypes.c_void_p(streams[r].cuda_stream)), "sgd_step_peer")
        torch.cuda.synchronize()
        w_exp, sh_exp = exp[s]
        for r in range(world):
            assert int(states[r][2]) == 0, "a barrier timed out"
            assert int(states[r][0]) == s + 1
            lo, hi, _ = peer.slice_bounds(n, world, r)
            assert torch.equal(masters[r][lo:hi], w_exp[lo:hi])               # own slice of the master
            assert torch.equal(shadows[r].view(torch.int16), sh_exp.view(torch.int16))   # every rank's whole operand copy
            if share_master:
                assert torch.equal(masters[r], w_exp)
    # the slices partition the tensor
    cover = sorted(peer.slice_bounds(n, world, r)[:2] for r in range(world))
    assert cover[0][0] == 0 and cover[-1][1] == n and all(a[1] == b[0] for a, b in zip(cover, cover[1:]))


@pytest.mark.parametrize("world,n", [(2, 3 * 1024 * 37 + 13), (3, 1024 * 50 + 3), (4, 1024 * 1024 + 1024), (8, 8 * 1024 * 9 + 77)])
def test_peer_sgd_copy_engine_transport_virtual_ranks(world, n):
    """transport="copy" of peer.PeerShardedSGD, replayed with raw pointers: push gradient slices into the owners' staging
    areas (da_peer_copy), all-local update kernel (DA_PEER_PUBLISH_BY_CALLER), push the bf16 slices, da_peer_publish_done."""
    g = torch.Generator(device=DEV).manual_seed(world * 77 + n % 89)
    w0 = torch.randn(n, device=DEV, generator=g)
    steps = 3
    grads_per_step = [[torch.randn(n, device=DEV, generator=g) for _ in range(world)] for _ in range(steps)]
    exp = _expected(w0, grads_per_step, world)
    bounds = [peer.slice_bounds(n, world, r) for r in range(world)]
    per = bounds[0][2]
    masters = [w0.clone() for _ in range(world)]
    grads = [torch.empty(n, device=DEV) for _ in range(world)]
    staging = [torch.full((world * per,), float("nan"), device=DEV) for _ in range(world)]
    shadows = [torch.zeros(n, device=DEV, dtype=torch.bfloat16) for _ in range(world)]
    flags = [torch.zeros(_lib.DA_PEER_FLAG_INTS, dtype=torch.int32, device=DEV) for _ in range(world)]
    states = [torch.zeros(4, dtype=torch.int32, device=DEV) for _ in range(world)]
    moms = [torch.zeros(max(per, 8), device=DEV) for _ in range(world)]
    args = []
    for r in range(world):
        lo = bounds[r][0]
        gp = [grads[r].data_ptr() if q == r else staging[r].data_ptr() + 4 * (q * per - lo) for q in range(world)]
        sp = [shadows[r].data_ptr() if q == r else 0 for q in range(world)]
        args.append(peer.make_args(masters[r].data_ptr(), moms[r].data_ptr(), gp, sp, None, [t.data_ptr() for t in flags],
                                   states[r].data_ptr(), n, world, r))
    streams = [torch.cuda.Stream() for _ in range(world)]
    for s in range(steps):
        for r in range(world):
            grads[r].copy_(grads_per_step[s][r])
        torch.cuda.synchronize()
        for r in range(world):
            st = ctypes.c_void_p(streams[r].cuda_stream)
            for q in range(world):
                if q != r:
                    lo, hi, _ = bounds[q]
                    check(lib.da_peer_copy(staging[q].data_ptr() + 4 * r * per, grads[r].data_ptr() + 4 * lo, 4 * (hi - lo), st), "copy")
        for r in range(world):
            st = ctypes.c_void_p(streams[r].cuda_stream)
            lo, hi, _ = bounds[r]
            check(lib.da_sgd_step_peer(ctypes.byref(args[r]), LR, MU, WD, int(s == 0), 8, _lib.DA_PEER_PUBLISH_BY_CALLER, st), "step")
            for q in range(world):
                if q != r:
                    check(lib.da_peer_copy(shadows[q].data_ptr() + 2 * lo, shadows[r].data_ptr() + 2 * lo, 2 * (hi - lo), st), "copy")
            check(lib.da_peer_publish_done(ctypes.byref(args[r]), st), "publish_done")
        torch.cuda.synchronize()
        w_exp, sh_exp = exp[s]
        for r in range(world):
            lo, hi, _ = bounds[r]
            assert int(states[r][2]) == 0 and int(states[r][0]) == s + 1
            assert torch.equal(masters[r][lo:hi], w_exp[lo:hi])
            assert torch.equal(shadows[r].view(torch.int16), sh_exp.view(torch.int16))
            assert int(flags[r][_lib.DA_MAX_PEERS:_lib.DA_MAX_PEERS + world].min()) == s + 1     # every rank raised done


@pytest.mark.parametrize("world,n", [(2, 1024 * 300 + 5), (4, 1024 * 1024), (8, 8 * 1024 * 33 + 9)])
def test_peer_sgd_deferred_publish_virtual_ranks(world, n):
    """The deferred publish of peer.PeerShardedSGD: the pushes of step s and da_peer_signal_done run on a SECOND stream per
    rank, concurrently with the beginning of step s+1, whose first read of the operand copy sits behind da_peer_wait_done;
    the update kernel of step s+1 waits for the rank's own publish s (local_state[3]).  No host synchronisation between
    steps; the operand copy every rank sees at the start of each step is snapshotted in stream order and compared bit for bit."""
    g = torch.Generator(device=DEV).manual_seed(world * 131 + n % 83)
    w0 = torch.randn(n, device=DEV, generator=g)
    steps = 4
    grads_per_step = [[torch.randn(n, device=DEV, generator=g) for _ in range(world)] for _ in range(steps)]
    exp = _expected(w0, grads_per_step, world)
    bounds = [peer.slice_bounds(n, world, r) for r in range(world)]
    per = bounds[0][2]
    masters = [w0.clone() for _ in range(world)]
    staging = [torch.zeros(world * per, device=DEV) for _ in range(world)]
    shadows = [torch.zeros(n, device=DEV, dtype=torch.bfloat16) for _ in range(world)]
    snaps = [[torch.zeros(n, device=DEV, dtype=torch.bfloat16) for _ in range(steps)] for _ in range(world)]
    flags = [torch.zeros(_lib.DA_PEER_FLAG_INTS, dtype=torch.int32, device=DEV) for _ in range(world)]
    states = [torch.zeros(4, dtype=torch.int32, device=DEV) for _ in range(world)]
    moms = [torch.zeros(max(per, 8), device=DEV) for _ in range(world)]
    grads = [[grads_per_step[s][r] for s in range(steps)] for r in range(world)]     # one buffer per step: no host sync needed
    args = [[None] * steps for _ in range(world)]
    for r in range(world):
        lo = bounds[r][0]
        for s in range(steps):
            gp = [grads[r][s].data_ptr() if q == r else staging[r].data_ptr() + 4 * (q * per - lo) for q in range(world)]
            sp = [shadows[r].data_ptr() if q == r else 0 for q in range(world)]
            args[r][s] = peer.make_args(masters[r].data_ptr(), moms[r].data_ptr(), gp, sp, None, [t.data_ptr() for t in flags],
                                        states[r].data_ptr(), n, world, r)
    main = [torch.cuda.Stream() for _ in range(world)]
    side = [torch.cuda.Stream() for _ in range(world)]
    torch.cuda.synchronize()
    for s in range(steps):
        for r in range(world):
            st = ctypes.c_void_p(main[r].cuda_stream)
            check(lib.da_peer_wait_done(ctypes.byref(args[r][s]), st), "wait_done")      # peers' pushes of step s-1 have landed
            with torch.cuda.stream(main[r]):
                snaps[r][s].copy_(shadows[r])                                            # the "forward" reads the operand copy
            for q in range(world):                                                       # gradient slices to their owners
                if q != r:
                    lo, hi, _ = bounds[q]
                    check(lib.da_peer_copy(staging[q].data_ptr() + 4 * r * per, grads[r][s].data_ptr() + 4 * lo, 4 * (hi - lo), st), "copy")
        for r in range(world):
            st = ctypes.c_void_p(main[r].cuda_stream)
            check(lib.da_sgd_step_peer(ctypes.byref(args[r][s]), LR, MU, WD, int(s == 0), 4, _lib.DA_PEER_PUBLISH_BY_CALLER, st), "step")
            ev = torch.cuda.Event()
            ev.record(main[r])
            side[r].wait_event(ev)
            sst = ctypes.c_void_p(side[r].cuda_stream)
            lo, hi, _ = bounds[r]
            for q in range(world):
                if q != r:
                    check(lib.da_peer_copy(shadows[q].data_ptr() + 2 * lo, shadows[r].data_ptr() + 2 * lo, 2 * (hi - lo), sst), "copy")
            check(lib.da_peer_signal_done(ctypes.byref(args[r][s]), sst), "signal_done")
    torch.cuda.synchronize()
    for r in range(world):
        assert int(states[r][2]) == 0, f"time-out code {int(states[r][2])}"
        assert int(states[r][0]) == steps and int(states[r][3]) == steps
        assert torch.equal(shadows[r].view(torch.int16), exp[-1][1].view(torch.int16))
        for s in range(1, steps):
            assert torch.equal(snaps[r][s].view(torch.int16), exp[s - 1][1].view(torch.int16)), f"rank {r} read a stale copy at step {s}"


def test_peer_sgd_rejects_bad_arguments():
    t = torch.zeros(1024, device=DEV)
    f = torch.zeros(_lib.DA_PEER_FLAG_INTS, dtype=torch.int32, device=DEV)
    st = torch.zeros(4, dtype=torch.int32, device=DEV)
    a = peer.make_args(t.data_ptr(), t.data_ptr(), [t.data_ptr()], [0], None, [f.data_ptr()], st.data_ptr(), 1024, 1, 0)
    a.world = 9
    assert lib.da_sgd_step_peer(ctypes.byref(a), LR, MU, WD, 0, 8, 0, None) != 0
    assert "world" in _lib.last_error()
    a.world = 1
    assert lib.da_sgd_step_peer(ctypes.byref(a), LR, MU, WD, 0, 8, 7, None) != 0
    assert "publish_mode" in _lib.last_error()
    a.w = t.data_ptr() + 4
    assert lib.da_sgd_step_peer(ctypes.byref(a), LR, MU, WD, 0, 8, 0, None) != 0
    assert "aligned" in _lib.last_error()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (CUDA IPC between processes)")
def test_peer_sharded_sgd_two_processes():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(root, "tools", "peer_check.py")]
    out = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "PEER_CHECK_OK" in out.stdout
