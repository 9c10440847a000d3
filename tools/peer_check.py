"""torchrun --nproc-per-node N tools/peer_check.py [--time]
Multi-process check of peer.PeerShardedSGD (CUDA IPC transport + da_sgd_step_peer) against the path it replaces:
NCCL all-reduce (AVG) of the gradient followed by da_sgd_step on every rank.  Prints PEER_CHECK_OK on rank 0.
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
    exact = world == 2          # NCCL sums in its own order for more than two ranks

    # ---- A: a layer driven through functional.dense_layer, managed vs all-reduce + FusedSGD ---------------------
    torch.manual_seed(0)
    lin_a = torch.nn.Linear(4096, 1024).to(dev)
    lin_b = torch.nn.Linear(4096, 1024).to(dev)
    lin_b.load_state_dict(lin_a.state_dict())
    transport = "stores" if "--stores" in sys.argv else "copy"
    popt = peer.PeerShardedSGD([lin_a.weight], lr=LR, momentum=MU, weight_decay=WD, transport=transport,
                               share_master="--share-master" in sys.argv)
    opt_a = optim.FusedSGD([lin_a.bias], lr=LR, momentum=MU, weight_decay=WD)
    opt_b = optim.FusedSGD(list(lin_b.parameters()), lr=LR, momentum=MU, weight_decay=WD)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    for step in range(4):
        x = torch.randn(256, 1, 1, 4096, device=dev, generator=g).to(torch.bfloat16)
        t = torch.randn(256, 1, 1, 1024, device=dev, generator=g)
        for lin in (lin_a, lin_b):
            y = F_.dense_layer(x, lin.weight, None, lin.bias, relu=True)
            (y.float() * t).sum().backward()
        assert lin_a.weight.grad is None
        for p in (lin_a.bias, lin_b.bias, lin_b.weight):
            dist.all_reduce(p.grad, op=dist.ReduceOp.AVG)
        opt_a.step(); opt_a.zero_grad()
        opt_b.step(); opt_b.zero_grad()
        popt.join()
        torch.cuda.synchronize()
        sa, sb = F_.bf16_shadow(lin_a.weight), F_.bf16_shadow(lin_b.weight)
        lo, hi, _ = peer.slice_bounds(lin_a.weight.numel(), world, rank)
        wa, wb = lin_a.weight.data.view(-1)[lo:hi], lin_b.weight.data.view(-1)[lo:hi]
        if exact:
            assert torch.equal(sa.view(torch.int16), sb.view(torch.int16)), f"step {step}: operand copies differ"
            assert torch.equal(wa, wb), f"step {step}: master slice differs"
        else:
            assert float((wa - wb).abs().max()) <= 1e-6 * float(wb.abs().max())
            assert float((sa.float() - sb.float()).abs().max()) <= 2 ** -7 * float(sb.float().abs().max())
        if exact:
            assert torch.equal(lin_a.bias.data, lin_b.bias.data)
        else:     # the two weight trajectories differ in the last bits, so do the activations behind them
            assert float((lin_a.bias.data - lin_b.bias.data).abs().max()) <= 1e-4 * float(lin_b.bias.data.abs().max())
    popt.check_errors()
    popt.gather_master()
    torch.cuda.synchronize()
    if exact:
        assert torch.equal(lin_a.weight.data, lin_b.weight.data), "gathered master differs"
    # every rank holds the same operand copy
    sa = F_.bf16_shadow(lin_a.weight).view(torch.int16).to(torch.int32)
    ref = sa.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(sa, ref)

    # ---- B: CUDA-graph replay of the peer kernel (device-resident epochs) ------------------------------------
    p = torch.nn.Parameter(torch.randn(1 << 22, device=dev, generator=torch.Generator(device=dev).manual_seed(5)))
    q = p.detach().clone()
    qbuf, qsh = torch.zeros_like(q), torch.empty_like(q, dtype=torch.bfloat16)
    popt2 = peer.PeerShardedSGD([p], lr=LR, momentum=MU, weight_decay=WD, transport=transport)
    mw = F_.MANAGED_WGRAD[id(p)]
    gbuf, wgrad_done, done = mw.grad, mw.after_wgrad, mw.layer_done
    src = torch.randn(1 << 22, device=dev, generator=g)

    def one_step():
        gbuf.copy_(src)
        wgrad_done()
        done()
        popt2.join()

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        one_step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        one_step()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    popt2.check_errors()
    gm = src.clone()
    dist.all_reduce(gm, op=dist.ReduceOp.AVG)
    for s in range(4):
        check(lib.da_sgd_step(F_._ptr(q), F_._ptr(gm), F_._ptr(qbuf), q.numel(), LR, MU, WD, int(s == 0), F_._ptr(qsh), None), "sgd")
    torch.cuda.synchronize()
    sh = F_.bf16_shadow(p).view(-1)
    if exact:
        assert torch.equal(sh.view(torch.int16), qsh.view(torch.int16)), "graph replay: operand copy differs"
    else:
        assert float((sh.float() - qsh.float()).abs().max()) <= 2 ** -7 * float(qsh.float().abs().max())

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
