#!/usr/bin/env python
"""Training entry point with the argument surface of the reference's tools/DA_train.py
(config, --work-dir, --resume-from, --load-from, --cfg-options, --seed, --launcher), driving this repo's modules:

    Config.fromfile(config) [+ --cfg-options]  ->  build_detector(cfg.model)            (DA_train.py:187-189,265-269)
    BatchSchedulerSampler / DistributedBatchSchedulerSampler over (source, target)       (datasets/builder.py:156-168)
    detector.train_step(data, optimizer) -> {loss, log_vars, num_samples}                (detectors/base.py:221-252)
    optimizer from cfg.optimizer (SGD: lr, momentum, weight_decay)  ->  optim.FusedSGD    (apis/train.py:127)
    N GPUs: dist.OverlappedGradAllReduce (NCCL) for the gradients                         (apis/train.py:113-121)
    checkpoint.save_checkpoint every --checkpoint-interval iterations / load on resume   (DA_train.py:258-263)

The image pipelines of the reference (decoding, resize, flip, Cityscapes annotations) are out of scope for this repo
(DESIGN.md 7), and no dataset exists in the build environment: `--synthetic N` trains on N seeded synthetic
source and N target images of `--img-size`.  A real loader only has to yield the same dicts
(img, img_metas, gt_bboxes, gt_labels, gt_da) in the sampler's order.

    python tools/DA_train.py tests/fixtures/cfg/experiment.py --synthetic 8 --img-size 128x192 --iters 4 --work-dir /tmp/da
    python -m torch.distributed.run --nproc-per-node 2 tools/DA_train.py <config> --launcher pytorch --synthetic 8 ...
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
from torch.utils.data import ConcatDataset, Dataset  # noqa: E402


def parse_args():
    ap = argparse.ArgumentParser(description="Train a DA detector")
    ap.add_argument("config")
    ap.add_argument("--work-dir", default=None)
    ap.add_argument("--resume-from", default=None, help="checkpoint to resume from (weights, optimizer momentum, iteration)")
    ap.add_argument("--load-from", default=None, help="checkpoint to initialise the weights from")
    ap.add_argument("--cfg-options", nargs="+", default=None, help="a.b.c=value overrides of the config")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--launcher", choices=["none", "pytorch"], default="none")
    ap.add_argument("--synthetic", type=int, default=0, help="number of synthetic images per domain")
    ap.add_argument("--img-size", default="256x512", help="HxW of the synthetic images")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--samples-per-gpu", type=int, default=2, help="images per GPU and iteration: half source, half target")
    ap.add_argument("--checkpoint-interval", type=int, default=0, help="iterations between checkpoints (0: only at the end)")
    ap.add_argument("--log-interval", type=int, default=1)
    return ap.parse_args()


class SyntheticDomainDataset(Dataset):
    """Seeded stand-in for one domain's DA_Dataset (mmdet/datasets/da_dataset.py): image, 1-3 boxes, labels."""

    def __init__(self, n, hw, num_classes, domain, seed):
        self.n, self.hw, self.num_classes, self.domain, self.seed = n, hw, max(1, num_classes), domain, seed

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed * 100003 + self.domain * 50021 + i)
        h, w = self.hw
        img = torch.randn(3, h, w, generator=g)
        k = int(torch.randint(1, 4, (1,), generator=g))
        cx, cy = torch.rand(k, generator=g) * w, torch.rand(k, generator=g) * h
        bw, bh = 16 + torch.rand(k, generator=g) * (w / 2), 16 + torch.rand(k, generator=g) * (h / 2)
        boxes = torch.stack([(cx - bw / 2).clamp(0, w - 2), (cy - bh / 2).clamp(0, h - 2), (cx + bw / 2).clamp(2, w), (cy + bh / 2).clamp(2, h)], 1)
        labels = torch.randint(0, self.num_classes, (k,), generator=g)
        meta = dict(img_shape=(h, w, 3), pad_shape=(h, w, 3), ori_shape=(h, w, 3), scale_factor=1.0, flip=False,
                    filename=f"synthetic_{'source' if self.domain == 0 else 'target'}_{i}")
        return dict(img=img, img_metas=meta, gt_bboxes=boxes, gt_labels=labels, gt_da=self.domain)


def collate(samples, device):
    """What mmcv's collate + scatter hand to train_step: a batch tensor and per-image lists."""
    return dict(img=torch.stack([s["img"] for s in samples]).to(device, non_blocking=True),
                img_metas=[s["img_metas"] for s in samples],
                gt_bboxes=[s["gt_bboxes"].to(device) for s in samples],
                gt_labels=[s["gt_labels"].to(device) for s in samples],
                gt_da=[s["gt_da"] for s in samples])


def main():
    args = parse_args()
    import unsupervised_domain_adaptation_object_detection_implementation_b200 as uda
    from unsupervised_domain_adaptation_object_detection_implementation_b200 import checkpoint, data, dist as ddist, optim

    rank, local, world = ddist.init_from_env("nccl") if args.launcher == "pytorch" else (0, 0, 1)
    if not torch.cuda.is_available():
        raise SystemExit("DA_train.py needs a CUDA device: the DA path has no CPU implementation")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    torch.manual_seed(args.seed)

    cfg = uda.Config.fromfile(args.config)
    if args.cfg_options:
        cfg.merge_from_dict(uda.Config.parse_cfg_options(args.cfg_options))
    work_dir = args.work_dir or os.path.join("work_dirs", os.path.splitext(os.path.basename(args.config))[0])
    model = uda.build_detector(cfg.model).to(dev).train()

    if not args.synthetic:
        raise SystemExit("no dataset pipeline in this repo (DESIGN.md 7): run with --synthetic N, or feed train_step() from your "
                         "own loader in data.BatchSchedulerSampler order")
    hw = tuple(int(v) for v in args.img_size.lower().split("x"))
    ncls = int(cfg.model.roi_head.bbox_head.get("num_classes", 1))
    dataset = ConcatDataset([SyntheticDomainDataset(args.synthetic, hw, ncls, d, args.seed) for d in (0, 1)])
    sampler = data.DistributedBatchSchedulerSampler(dataset, args.samples_per_gpu, world, rank, seed=args.seed)

    ocfg = dict(cfg.get("optimizer", dict(type="SGD", lr=0.001, momentum=0.9, weight_decay=0.0005)))
    if ocfg.pop("type", "SGD") != "SGD":
        raise SystemExit("the DA configs train with SGD (da_configs/faster_rcnn/*.py); other optimizers are not wired")
    unused = model.unused_parameters() if hasattr(model, "unused_parameters") else []
    params = ddist.trainable_parameters(model, unused)
    optimizer = optim.FusedSGD(params, lr=ocfg.get("lr", 1e-3), momentum=ocfg.get("momentum", 0.9), weight_decay=ocfg.get("weight_decay", 0.0))
    reducer = ddist.OverlappedGradAllReduce(params) if world > 1 else None
    schedule = optim.StepLrSchedule(optimizer.lr, cfg.get("lr_config"))      # step decay + linear warm-up (schedule_1x.py)

    start_epoch = 0
    start_iter, meta = 0, dict(seed=args.seed, config=os.path.abspath(args.config), exp_name=os.path.basename(args.config))
    if args.resume_from or args.load_from:
        ck = checkpoint.load_checkpoint(model, args.resume_from or args.load_from, strict=False)
        rep = ck["load_report"]
        if rank == 0 and (rep["missing_keys"] or rep["unexpected_keys"]):
            print(f"load: missing {rep['missing_keys'][:5]} unexpected {rep['unexpected_keys'][:5]}", flush=True)
        if args.resume_from:
            start_iter = int(ck.get("meta", {}).get("iter", 0))
            start_epoch = int(ck.get("meta", {}).get("epoch", 0))
            osd = ck.get("optimizer", {})
            if "momentum" in osd:                     # checkpoints of round 1: a bare list of buffers in parameter order
                osd = {"state": {i: {"momentum_buffer": b} for i, b in enumerate(osd["momentum"]) if b.numel()}, "param_groups": []}
            optimizer.load_state_dict(osd)            # torch.optim.SGD layout (FusedSGD.state_dict)
            optimizer.steps = start_iter

    def save(it, epoch):
        path = os.path.join(work_dir, f"iter_{it}.pth")
        checkpoint.save_checkpoint(model, path, optimizer=optimizer, meta=dict(meta, iter=it, epoch=epoch), rank=rank)
        if rank == 0:
            print(f"checkpoint: {path}", flush=True)

    it, t0 = start_iter, time.time()
    epoch = start_epoch                                   # the sampler's seeded schedule resumes in the epoch it stopped in
    while it < args.iters:
        sampler.set_epoch(epoch)
        order = list(iter(sampler))
        for k in range(0, len(order), args.samples_per_gpu):
            if it >= args.iters:
                break
            batch = collate([dataset[i] for i in order[k:k + args.samples_per_gpu]], dev)
            optimizer.lr = schedule(it, epoch)          # eager loop: the kernels take lr as a launch argument
            out = model.train_step(batch, optimizer)
            out["loss"].backward()
            if reducer is not None:
                reducer()
            optimizer.step()
            optimizer.zero_grad(set_to_none=True)
            it += 1
            if rank == 0 and it % args.log_interval == 0:
                lv = out["log_vars"]
                da = {k: round(float(v), 5) for k, v in lv.items() if "da_loss" in k or "consistency" in k or "patch" in k}
                print(f"iter {it}/{args.iters} lr {optimizer.lr:.3e} loss {float(lv['loss']):.4f} DA {da} domains {batch['gt_da']} "
                      f"({(time.time() - t0) / max(1, it - start_iter):.2f} s/iter)", flush=True)
            if args.checkpoint_interval and it % args.checkpoint_interval == 0:
                save(it, epoch)
        epoch += 1
    save(it, epoch)
    if world > 1:
        torch.distributed.barrier()
        torch.cuda.synchronize()
        os._exit(0)


if __name__ == "__main__":
    main()
