"""torchrun --nproc-per-node N tools/peer_check.py [--time]
Multi-process check of peer.PeerShardedSGD (CUDA IPC transport + da_sgd_step_peer) against the math it replaces:
mean of the per-rank gradients (summed in rank order) followed by da_sgd_step on every rank, bit for bit.  Prints PEER_CHECK_OK on rank 0.
With --time it also times the fused kernel on the FC1-sized tensor (1024 x 100352) next to all-reduce + SGD."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from unsupervised_domain_adaptation_object_detection_implementation_b200 import dist as ddist, functional as F_, optim, peer  # noqa: E402
from unsupervised_domain_adaptation_object_detection_implementation_b200._lib import check, lib  # noqa: E402

LR, MU, WD = 0.05, 0.9, 5e-4


def main():
    rank, local, world = ddist.init_from_env("nccl")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)

    def rank_order_mean(t):
        """Mean over ranks summed in rank order 0..N-1 (what the peer kernel computes), identical bits on every rank."""
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t.contiguous())
        acc = parts[0].clone()
        for q in parts[1:]:
            acc = acc + q
        return acc * torch.tensor(1.0 / world, dtype=torch.float32, device=dev)

    # ---- A: a layer driven through functional.dense_layer; reference = rank-order mean of the per-rank gradients
    #         (read back from the managed gradient buffer) followed by da_sgd_step: bit-exact for every world size
    torch.manual_seed(0)
    lin_a = torch.nn.Linear(4096, 1024).to(dev)
    transport = "stores" if "--stores" in sys.argv else "copy"
    popt = peer.PeerShardedSGD([lin_a.weight], lr=LR, momentum=MU, weight_decay=WD, transport=transport,
                               share_master="--share-master" in sys.argv)
    n = lin_a.weight.numel()
    w_ref = lin_a.weight.data.detach().clone().view(-1)
    buf_ref = torch.zeros_like(w_ref)
    sh_ref = torch.zeros(n, dtype=torch.bfloat16, device=dev)
    mw = F_.MANAGED_WGRAD[id(lin_a.weight)]
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    lo, hi, _ = peer.slice_bounds(n, world, rank)
    for step in range(4):
        x = torch.randn(256, 1, 1, 4096, device=dev, generator=g).to(torch.bfloat16)
        t = torch.randn(256, 1, 1, 1024, device=dev, generator=g)
        y = F_.dense_layer(x, lin_a.weight, None, lin_a.bias, relu=True)
        (y.float() * t).sum().backward()
        assert lin_a.weight.grad is None
        lin_a.bias.grad = None
        popt.join()
        torch.cuda.synchronize()
        gm = rank_order_mean(mw.grad)
        check(lib.da_sgd_step(F_._ptr(w_ref), F_._ptr(gm), F_._ptr(buf_ref), n, LR, MU, WD, int(step == 0), F_._ptr(sh_ref), None), "sgd")
        torch.cuda.synchronize()
        assert torch.equal(lin_a.weight.data.view(-1)[lo:hi], w_ref[lo:hi]), f"step {step}: master slice differs"
        assert torch.equal(F_.bf16_shadow(lin_a.weight).view(-1).view(torch.int16), sh_ref.view(torch.int16)), f"step {step}: operand copy differs"
    popt.check_errors()
    popt.gather_master()
    torch.cuda.synchronize()
    assert torch.equal(lin_a.weight.data.view(-1), w_ref), "gathered master differs"

    # ---- B: CUDA-graph replay of the peer kernel (device-resident epochs), publish inside the step and deferred ----
    for deferred in (False, True):
        p = torch.nn.Parameter(torch.randn(1 << 22, device=dev, generator=torch.Generator(device=dev).manual_seed(5)))
        q = p.detach().clone()
        qbuf, qsh = torch.zeros_like(q), torch.empty_like(q, dtype=torch.bfloat16)
        popt2 = peer.PeerShardedSGD([p], lr=LR, momentum=MU, weight_decay=WD, transport=transport, deferred_publish=deferred)
        mw = F_.MANAGED_WGRAD[id(p)]
        gbuf, wgrad_done, done = mw.grad, mw.after_wgrad, mw.layer_done
        src = torch.randn(1 << 22, device=dev, generator=g)
        probe = torch.zeros(8, device=dev)

        def one_step():
            mw.before_forward()                       # what functional.dense_layer does in front of the layer's forward
            probe.copy_(F_.bf16_shadow(p).view(-1)[:8].float())    # a "forward" that reads the operand copy
            gbuf.copy_(src)
            wgrad_done()
            done()
            popt2.join()

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            one_step()
            popt2.publish()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            one_step()
        for _ in range(3):
            graph.replay()
            popt2.publish()
        torch.cuda.synchronize()
        popt2.check_errors()
        gm = rank_order_mean(src)
        for s in range(4):
            check(lib.da_sgd_step(F_._ptr(q), F_._ptr(gm), F_._ptr(qbuf), q.numel(), LR, MU, WD, int(s == 0), F_._ptr(qsh), None), "sgd")
        torch.cuda.synchronize()
        sh = F_.bf16_shadow(p).view(-1)
        assert torch.equal(sh.view(torch.int16), qsh.view(torch.int16)), f"graph replay (deferred={deferred}): operand copy differs"
        del F_.MANAGED_WGRAD[id(p)]

    # ---- C: timing at the FC1 size ----------------------------------------------------------------------------
    if "--time" in sys.argv:
        n = 1024 * 100352
        big = torch.nn.Parameter(torch.zeros(n, device=dev))
        for transport, ctas in (("copy", 0), ("copy", 592), ("stores", 296)):
            po = peer.PeerShardedSGD([big], lr=LR, momentum=MU, weight_decay=WD, max_ctas=ctas, transport=transport)
            wgrad_big, done_big = F_.MANAGED_WGRAD[id(big)].after_wgrad, F_.MANAGED_WGRAD[id(big)].layer_done
            ts = []
            for it in range(6):
                dist.barrier()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                wgrad_big()
                done_big()
                po.join()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            po.check_errors()
            t = ddist.max_over_ranks(sorted(ts[1:])[len(ts[1:]) // 2], dev)
            if rank == 0:
                lo, hi, _ = peer.slice_bounds(n, world, 0)
                remote_in = (world - 1) * (hi - lo) * 4
                print(f"PEER_TIME world={world} transport={transport} ctas={ctas} ms={t:.4f} remote_in_GBps={remote_in / t / 1e6:.1f} "
                      f"remote_out_GBps={(world - 1) * (hi - lo) * 2 / t / 1e6:.1f}", flush=True)
            del F_.MANAGED_WGRAD[id(big)]
        grad = torch.zeros(n, device=dev)
        buf = torch.zeros(n, device=dev)
        sh = torch.zeros(n, device=dev, dtype=torch.bfloat16)
        ts = []
        for it in range(6):
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dist.all_reduce(grad, op=dist.ReduceOp.AVG)
            check(lib.da_sgd_step(F_._ptr(big.data), F_._ptr(grad), F_._ptr(buf), n, LR, MU, WD, 0, F_._ptr(sh), F_._stream()), "sgd")
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = ddist.max_over_ranks(sorted(ts[1:])[len(ts[1:]) // 2], dev)
        if rank == 0:
            print(f"NCCL_TIME world={world} allreduce+sgd ms={t:.4f}", flush=True)

    dist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        print("PEER_CHECK_OK", flush=True)
    os._exit(0)


if __name__ == "__main__":
    main()
