"""Multi-process logic of the data-parallel path on CPU (gloo, world_size 2): pair sharding, the
flat-buffer gradient all-reduce and the max-over-ranks timing reduction used by bench.py."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from unsupervised_domain_adaptation_object_detection_implementation_b200 import dist as ddist


def test_shard_pairs_is_a_balanced_partition():
    for pairs in (1, 2, 7, 8, 16, 17):
        for world in (1, 2, 3, 4, 8):
            parts = [ddist.shard_pairs(pairs, r, world) for r in range(world)]
            flat = [p for part in parts for p in part]
            assert flat == list(range(pairs))                      # complete, ordered, disjoint
            sizes = [len(p) for p in parts]
            assert max(sizes) - min(sizes) <= 1                    # balanced; a pair is never split


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, l, w = ddist.init_from_env("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Linear(16, 4), torch.nn.Linear(4, 3))
    unused = list(model[2].parameters())                            # never receive a gradient (Q9/Q10 analogue)
    params = ddist.trainable_parameters(model, unused)
    assert len(params) == 4
    x = torch.full((5, 8), float(rank + 1))
    model[1](model[0](x)).sum().backward()
    local = [p.grad.clone() for p in params]
    params[3].grad = None                       # no gradient on ANY rank: skipped, stays None (as on one GPU: no decay/momentum)
    if rank == 1:
        params[2].grad = None                   # missing on ONE rank only: that rank contributes zeros and receives the mean
        local[2] = torch.zeros_like(local[2])
    red = ddist.FlatGradAllReduce(params)
    red()
    assert params[3].grad is None and params[2].grad is not None
    kept = [p.grad.clone() for p in params[:3]]
    for p, g in zip(params[:3], local[:3]):     # second step with the same pattern: the cached agreement is reused
        p.grad = g.clone() if not (rank == 1 and p is params[2]) else None
    red()
    assert all(torch.equal(a, p.grad) for a, p in zip(kept, params[:3]))
    gathered = [torch.zeros_like(torch.cat([g.reshape(-1) for g in local])) for _ in range(world)]
    dist.all_gather(gathered, torch.cat([g.reshape(-1) for g in local]))
    mine = torch.cat([p.grad.reshape(-1) for p in params[:3]] + [torch.zeros(params[3].numel())])
    tmax = ddist.max_over_ranks(10.0 * (rank + 1), torch.device("cpu"))
    if rank == 0:
        torch.save({"mine": mine, "gathered": gathered, "tmax": tmax, "n3": params[3].numel()}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_flat_grad_allreduce_world2_gloo(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    rec = torch.load(out)
    expect = (rec["gathered"][0] + rec["gathered"][1]) / 2
    n3 = rec["n3"]
    expect[-n3:] = 0.0                                               # both ranks dropped that gradient
    assert torch.allclose(rec["mine"], expect, atol=1e-6)
    assert rec["tmax"] == 20.0


def test_peer_slice_bounds_partition_the_tensor():
    """Slice ownership of peer.PeerShardedSGD (same rule as the kernel): contiguous, disjoint, complete, multiples of 1024
    elements (16-byte aligned for fp32 and bf16), identical `per` on every rank."""
    from unsupervised_domain_adaptation_object_detection_implementation_b200 import peer
    for n in (1, 1023, 1024, 1025, 8192 + 5, 1024 * 100352, 3 * 1024 * 37 + 13):
        for world in (1, 2, 3, 4, 5, 8):
            b = [peer.slice_bounds(n, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(x[1] == y[0] for x, y in zip(b, b[1:]))
            assert len({x[2] for x in b}) == 1 and b[0][2] % 1024 == 0
            assert all(x[0] % 1024 == 0 or x[0] == n for x in b)
            assert all(0 <= x[1] - x[0] <= x[2] for x in b)


def _sampler_ckpt_worker(rank, world, port, outdir):
    """world_size-2 gloo: each rank draws its sample stream from DistributedBatchSchedulerSampler (rank/world taken from the
    process group) and all ranks call checkpoint.save_checkpoint; only rank 0 may write."""
    from torch.utils.data import ConcatDataset, TensorDataset
    from unsupervised_domain_adaptation_object_detection_implementation_b200 import checkpoint, data
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, _, w = ddist.init_from_env("gloo")
    ds = ConcatDataset([TensorDataset(torch.zeros(9)), TensorDataset(torch.zeros(5))])
    smp = data.DistributedBatchSchedulerSampler(ds, samples_per_gpu=2, num_replicas=w, rank=r, seed=3)
    mine = torch.tensor(list(iter(smp)), dtype=torch.long)
    streams = [torch.zeros_like(mine) for _ in range(w)]
    dist.all_gather(streams, mine)                                   # equal length on every rank, or this call fails
    torch.manual_seed(0)
    model = torch.nn.Linear(4, 3)
    path = os.path.join(outdir, "ck.pth")
    checkpoint.save_checkpoint(model, path, meta={"iter": 7}, rank=r)
    dist.barrier()
    if r == 1:
        torch.save({"streams": streams, "schedule": smp.global_schedule(), "exists": os.path.exists(path)}, os.path.join(outdir, "r1.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_distributed_pair_sampler_and_rank0_checkpoint_world2_gloo(tmp_path):
    mp.spawn(_sampler_ckpt_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    rec = torch.load(str(tmp_path / "r1.pt"), weights_only=False)
    s0, s1 = rec["streams"][0].tolist(), rec["streams"][1].tolist()
    pairs = [tuple(s[k:k + 2]) for s in (s0, s1) for k in range(0, len(s), 2)]
    assert sorted(pairs) == sorted(tuple(b) for b in rec["schedule"])           # the ranks partition the epoch's pairs
    assert all(a < 9 <= b for a, b in pairs)                                    # (source, target), never split
    assert rec["exists"]
    ck = torch.load(str(tmp_path / "ck.pth"), weights_only=False)
    assert ck["meta"]["iter"] == 7 and set(ck["state_dict"]) == {"weight", "bias"}
