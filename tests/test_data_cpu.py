"""Pair-aware sampling and checkpoint I/O, CPU only."""
import json
import os

import torch
import torch.nn as nn
from torch.utils.data import ConcatDataset, TensorDataset

from unsupervised_domain_adaptation_object_detection_implementation_b200 import checkpoint, da_heads, data, dist as ddist

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _concat(sizes):
    return ConcatDataset([TensorDataset(torch.zeros(n)) for n in sizes])


def test_batch_scheduler_sampler_reproduces_the_reference_index_stream():
    """Golden: the reference's own BatchSchedulerSampler (mmdet/datasets/samplers/batch_sampler.py, loaded in place by
    oracle/make_golden.sampler_cases) after torch.manual_seed(seed)."""
    cases = json.load(open(os.path.join(GOLDEN, "batch_scheduler_sampler.json")))
    assert len(cases) >= 4
    for c in cases:
        smp = data.BatchSchedulerSampler(_concat(c["sizes"]), samples_per_gpu=c["samples_per_gpu"])
        assert len(smp) == c["len"]
        torch.manual_seed(c["seed"])
        assert list(iter(smp)) == c["indices"], c["sizes"]


def test_every_minibatch_is_source_then_target():
    ds = _concat((11, 5))
    smp = data.BatchSchedulerSampler(ds, samples_per_gpu=4)
    torch.manual_seed(0)
    idx = list(iter(smp))
    assert len(idx) % 4 == 0
    for k in range(0, len(idx), 4):
        assert [data.domain_of(ds, i) for i in idx[k:k + 4]] == [0, 0, 1, 1]
    assert set(i for i in idx if i < 11) == set(range(11))              # the larger dataset is covered once
    assert set(i - 11 for i in idx if i >= 11) == set(range(5))         # the smaller one is restarted until then


def test_distributed_sampler_deals_whole_pairs_and_equal_work():
    ds = _concat((13, 6))
    for world in (1, 2, 3, 4, 8):
        smps = [data.DistributedBatchSchedulerSampler(ds, 2, world, r, seed=5) for r in range(world)]
        streams = [list(iter(s)) for s in smps]
        assert len({len(s) for s in streams}) == 1 and all(len(s) == len(smp) for s, smp in zip(streams, smps))
        glob = smps[0].global_schedule()
        assert all(s.global_schedule() == glob for s in smps)           # identical schedule on every rank
        dealt = [tuple(st[k:k + 2]) for st in streams for k in range(0, len(st), 2)]
        assert sorted(dealt) == sorted(tuple(b) for b in glob)          # every mini-batch to exactly one rank
        for st in streams:                                              # a pair is never split
            assert all([data.domain_of(ds, i) for i in st[k:k + 2]] == [0, 1] for k in range(0, len(st), 2))
        smps[0].set_epoch(1)
        assert smps[0].global_schedule() != glob


def test_checkpoint_roundtrip_in_reference_layout(tmp_path):
    torch.manual_seed(0)
    model = nn.ModuleDict({"local_da": da_heads.InstanceAlignmentHead(), "da_head_top": da_heads.ImgAlignmentHead(64)})
    model.CLASSES = ("person", "car")
    path = str(tmp_path / "epoch_1.pth")
    ck = checkpoint.save_checkpoint(model, path, meta={"epoch": 1, "iter": 10})
    assert set(ck) == {"meta", "state_dict"} and ck["meta"]["CLASSES"] == ("person", "car") and ck["meta"]["epoch"] == 1
    raw = torch.load(path, map_location="cpu", weights_only=False)
    assert list(raw["state_dict"]) == list(model.state_dict())          # reference key names, same order
    assert all(not v.is_cuda for v in raw["state_dict"].values())
    other = nn.ModuleDict({"local_da": da_heads.InstanceAlignmentHead(), "da_head_top": da_heads.ImgAlignmentHead(64)})
    other["local_da"].fc1.weight._da_shadow = (0, torch.zeros(1))       # a stale cached operand copy
    rep = checkpoint.load_checkpoint(other, path, strict=True)
    assert rep["load_report"] == {"missing_keys": [], "unexpected_keys": []}
    assert not hasattr(other["local_da"].fc1.weight, "_da_shadow")
    for (k, a), (_, b) in zip(model.state_dict().items(), other.state_dict().items()):
        assert torch.equal(a, b), k


def test_load_checkpoint_strips_dataparallel_prefix_and_reports(tmp_path):
    head = da_heads.InstanceAlignmentHead_DAF()
    sd = {"module." + k: v for k, v in head.state_dict().items()}
    sd["module.extra.weight"] = torch.zeros(1)
    del sd["module.fc3.bias"]
    path = str(tmp_path / "dp.pth")
    torch.save({"state_dict": sd, "meta": {}}, path)
    fresh = da_heads.InstanceAlignmentHead_DAF()
    rep = checkpoint.load_checkpoint(fresh, path)["load_report"]
    assert rep["missing_keys"] == ["fc3.bias"] and rep["unexpected_keys"] == ["extra.weight"]
    assert torch.equal(fresh.fc1.weight, head.fc1.weight)
    torch.save(head.state_dict(), path)                                  # a bare state_dict is accepted too
    assert checkpoint.load_checkpoint(da_heads.InstanceAlignmentHead_DAF(), path, strict=True)["load_report"]["missing_keys"] == []


def test_save_checkpoint_writes_on_rank_zero_only(tmp_path):
    head = da_heads.InstanceAlignmentHead_DAF()
    assert checkpoint.save_checkpoint(head, str(tmp_path / "r1.pth"), rank=1) is None
    assert not os.path.exists(tmp_path / "r1.pth")
    assert ddist.shard_pairs(4, 1, 2) == [2, 3]


def test_step_lr_schedule_matches_mmcv_step_updater_with_linear_warmup():
    """optim.StepLrSchedule against mmcv's StepLrUpdaterHook + linear warm-up formulas for the DAF recipe
    (da_configs/faster_rcnn/faster_rcnn_r50_daf_c2f.py:13-18: warmup 500 iters from ratio 1e-4, step at epoch 9)."""
    from unsupervised_domain_adaptation_object_detection_implementation_b200 import optim
    cfg = dict(policy="step", warmup="linear", warmup_iters=500, warmup_ratio=0.0001, step=[9])
    sch = optim.StepLrSchedule(1e-3, cfg)
    assert abs(sch(0, 0) - 1e-3 * 1e-4) < 1e-12                       # k = (1 - 0/500) * (1 - ratio) -> lr * ratio
    assert abs(sch(250, 0) - 1e-3 * (1 - 0.5 * (1 - 1e-4))) < 1e-12
    assert sch(500, 0) == 1e-3 and sch(10_000, 8) == 1e-3
    assert abs(sch(10_000, 9) - 1e-4) < 1e-15 and abs(sch(10_000, 13) - 1e-4) < 1e-15
    two = optim.StepLrSchedule(0.02, dict(policy="step", step=[6, 8]))
    assert [round(two(1000, e), 8) for e in (0, 5, 6, 7, 8, 9)] == [0.02, 0.02, 0.002, 0.002, 0.0002, 0.0002]
