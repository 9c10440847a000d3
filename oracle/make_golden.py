"""ORACLE — test infrastructure only.  Generates tests/golden/*.pt in the BUILD CONTAINER.

Runs the reference's own classes (loaded in place from /root/reference by oracle/ref_loader.py) and
torchvision's C++ RoIAlign (the stand-in for mmcv-full 1.3.17's kernel) on seeded
inputs, and stores the small input/output vectors.  Parameters are NOT stored: they are a pure
function of (state_dict key, seed) — oracle/seeded.py — on both sides.

    python -m oracle.make_golden            # from the repo root
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader, seeded  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _summ(t):
    """Large tensors are stored as (norm, first 32, last 32, strided sample)."""
    t = t.detach().float().reshape(-1)
    if t.numel() <= 4096:
        return {"full": t.clone()}
    idx = torch.linspace(0, t.numel() - 1, 512).long()
    return {"norm": t.norm().item(), "head": t[:32].clone(), "tail": t[-32:].clone(), "idx": idx, "sample": t[idx].clone()}


def head_case(name, module, x, seed=0):
    module = module.float().eval()  # eval: dropout off, BN frozen (== training with norm_eval, Q9)
    seeded.fill_state_(module, seed, prefix=name + ".")
    x = x.clone().requires_grad_(True)
    out = module(x)
    outs = out if isinstance(out, tuple) else (out,)
    cot = [seeded.seeded_tensor(f"{name}.cot{i}", o.shape, seed) for i, o in enumerate(outs)]
    loss = sum((o * c).sum() for o, c in zip(outs, cot))
    loss.backward()
    rec = {"x": x.detach().clone(), "out": [o.detach().clone() for o in outs], "cot": cot,
           "dx": x.grad.detach().clone(), "seed": seed,
           "dparams": {k: _summ(p.grad) for k, p in module.named_parameters() if p.grad is not None},
           "no_grad_params": [k for k, p in module.named_parameters() if p.grad is None]}
    torch.save(rec, os.path.join(OUT, f"head_{name}.pt"))
    print(f"head_{name}: out {[tuple(o.shape) for o in outs]} |dx|={x.grad.norm():.4f}")


def roi_align_case():
    import torchvision  # noqa: F401
    N, C, H, W, stride = 2, 8, 20, 30, 16
    feat = seeded.seeded_tensor("roi.feat", (N, C, H, W), 0)
    rois = torch.cat([seeded.synthetic_rois(24, N, H * stride, W * stride, 0, 16.0, 300.0),
                      seeded.adversarial_rois(N, H * stride, W * stride)], 0)
    f = feat.clone().requires_grad_(True)
    out = torch.ops.torchvision.roi_align(f, rois, 1.0 / stride, 7, 7, 0, True)
    cot = seeded.seeded_tensor("roi.cot", out.shape, 0)
    (out * cot).sum().backward()
    # the sampling grid in fp32, same operation order as the kernels (no FMA on the CPU build)
    s = torch.tensor(1.0 / stride, dtype=torch.float32)
    x1, y1, x2, y2 = [rois[:, i] * s - 0.5 for i in (1, 2, 3, 4)]
    gh = torch.ceil((y2 - y1) / 7.0).int()
    gw = torch.ceil((x2 - x1) / 7.0).int()
    rec = {"feat": feat, "rois": rois, "stride": stride, "out": out.detach().clone(), "cot": cot,
           "dfeat": f.grad.detach().clone(), "grid": torch.stack([gh, gw], 1), "batch_idx": rois[:, 0].int()}
    # non-aligned, fixed sampling ratio variant (mmcv's legacy mode is reachable through the same module)
    out2 = torch.ops.torchvision.roi_align(feat, rois[:24], 1.0 / stride, 7, 7, 2, False)
    rec["out_legacy_sr2"] = out2.clone()
    torch.save(rec, os.path.join(OUT, "roi_align_torchvision.pt"))
    print("roi_align:", tuple(out.shape), "grid max", int(gh.max()), int(gw.max()))


def backbone_loss_cases(ns):
    """Run the reference backbones' forward_train on a tiny image: pins the loss tails L1/L2/L3 to
    the reference's own code (resnet_da_daf_org.py:796-824, resnet_da_cbam.py:934-993, resnet_da.py:821-850)."""
    torch.manual_seed(0)
    common = dict(depth=50, num_stages=4, strides=(1, 2, 2, 1), dilations=(1, 1, 1, 2), out_indices=(3,),
                  frozen_stages=1, norm_eval=True, style="pytorch")
    img = torch.randn(2, 3, 64, 96)
    gt = torch.tensor([0, 1])
    rec = {}
    # DAF-Org: L1 on ImgAlignmentHead output
    m = ns.daf_org.ResNet_DAF(**common)
    torch.nn.init.normal_(m.da_head_top.conv1.weight, 0, 0.05)
    torch.nn.init.normal_(m.da_head_top.conv2.weight, 0, 0.2)
    m.train()
    outs, loss, patch = m.forward_train(img, gt)
    rec["daf_org"] = {"patch_feat": patch.detach().clone(), "loss": loss.detach().clone(), "gt": gt}
    outs, loss_tt, _ = m.forward_train(img, torch.tensor([1, 1]))
    rec["daf_org_tt"] = {"loss": loss_tt.detach().clone(), "gt": torch.tensor([1, 1])}
    # MAF: L3 on sigmoid outputs of the three SRM heads (eval() for determinism: dropout off)
    m = ns.maf.ResNet_DA(**common)
    m.eval()
    feats = {}
    hooks = [getattr(m, n).register_forward_hook(lambda mod, i, o, n=n: feats.__setitem__(n, o.detach().clone()))
             for n in ("da_head_bottom", "da_head_mid", "da_head_top")]
    outs, losses = m.forward_train(img, gt)
    for h in hooks:
        h.remove()
    rec["maf"] = {"preds": [feats["da_head_bottom"], feats["da_head_mid"], feats["da_head_top"]],
                  "losses": losses.detach().clone(), "gt": gt}
    # CBAM: L3 on raw logits of the two Global heads + L2 on the Local head
    m = ns.cbam.ResNet_DA_CBAM(**common)
    torch.nn.init.normal_(m.local_da_head_bottom.conv3.weight, 0, 0.3)
    m.eval()
    feats = {}
    hooks = [getattr(m, n).register_forward_hook(lambda mod, i, o, n=n: feats.__setitem__(n, o.detach().clone()))
             for n in ("local_da_head_bottom", "da_head_mid", "da_head_top")]
    outs, glob, patch = m.forward_train(img, gt)
    for h in hooks:
        h.remove()
    rec["cbam"] = {"local_feat": feats["local_da_head_bottom"], "logits": [feats["da_head_mid"], feats["da_head_top"]],
                   "global_losses": glob.detach().clone(), "patch_loss": patch.detach().clone(), "gt": gt}
    torch.save(rec, os.path.join(OUT, "backbone_loss_tails.pt"))
    print("backbone loss tails: daf_org", float(rec["daf_org"]["loss"]), "maf", rec["maf"]["losses"].tolist(),
          "cbam", rec["cbam"]["global_losses"].tolist(), float(rec["cbam"]["patch_loss"]))


def focal_case(ns):
    u = seeded.seeded_tensor("focal.u", (40, 2), 0, scale=2.0).requires_grad_(True)
    labels = (seeded.seeded_tensor("focal.lab", (40,), 0) > 0).long()
    # FocalLoss.forward on CPU: one-hot with num_classes+1 then slice (focal_loss.py:166-168)
    t = torch.nn.functional.one_hot(labels, num_classes=3)[:, :2]
    loss = ns.focal.py_sigmoid_focal_loss(u, t, None, gamma=2.0, alpha=0.25, reduction="mean")
    loss.backward()
    torch.save({"u": u.detach().clone(), "labels": labels, "loss": loss.detach().clone(), "du": u.grad.clone()},
               os.path.join(OUT, "focal_loss.pt"))
    print("focal:", float(loss))


def bbox_head_loss_case(ns):
    """R3: the box-head loss of the source image as BBoxHead.loss computes it (mmdet/models/roi_heads/bbox_heads/bbox_head.py:
    256-315) with the reference's OWN loss classes: CrossEntropyLoss(use_sigmoid=True) reduced with avg_factor = number of
    sampled RoIs, SmoothL1Loss(beta=1) over the positives with avg_factor = number of sampled RoIs, top-1 accuracy."""
    R, C = 48, 3
    cls_score = seeded.seeded_tensor("bbox.cls", (R, C + 1), 0, scale=2.0).requires_grad_(True)
    bbox_pred = seeded.seeded_tensor("bbox.reg", (R, 4 * C), 0).requires_grad_(True)
    labels = (seeded.seeded_tensor("bbox.lab", (R,), 0, "uniform") * 2.7 - 1.2).clamp(0, C).long()    # classes 0..C-1, C = background
    labels[:4] = torch.tensor([0, 1, 2, C])
    targets = seeded.seeded_tensor("bbox.tgt", (R, 4), 0, scale=0.7)
    label_weights = torch.ones(R)
    pos = (labels >= 0) & (labels < C)
    bbox_weights = pos.float()[:, None].expand(R, 4).contiguous()
    avg_factor = max(float((label_weights > 0).sum()), 1.0)
    loss_cls = ns.ce.CrossEntropyLoss(use_sigmoid=True, loss_weight=1.0)(cls_score, labels, label_weights, avg_factor=avg_factor)
    acc = ns.accuracy.accuracy(cls_score, labels)
    pos_pred = bbox_pred.view(R, -1, 4)[pos, labels[pos]]
    loss_bbox = ns.smooth_l1.SmoothL1Loss(beta=1.0, loss_weight=1.0)(pos_pred, targets[pos], bbox_weights[pos], avg_factor=targets.size(0))
    (loss_cls + loss_bbox).backward()
    torch.save({"cls_score": cls_score.detach().clone(), "bbox_pred": bbox_pred.detach().clone(), "labels": labels, "targets": targets,
                "pos_mask": pos, "num_classes": C, "loss_cls": loss_cls.detach().clone(), "loss_bbox": loss_bbox.detach().clone(),
                "acc": acc.detach().clone().reshape(()), "dcls": cls_score.grad.clone(), "dreg": bbox_pred.grad.clone()},
               os.path.join(OUT, "bbox_head_loss.pt"))
    print(f"bbox_head_loss: cls {float(loss_cls):.6f} bbox {float(loss_bbox):.6f} acc {float(acc):.2f} ({int(pos.sum())} positives of {R})")


def group_loss_cases(ns):
    """L5: the reference's own group_local_da_loss methods (compiled in place, ref_loader.load_group_loss) on CPU.
    Inputs are a pure function of the case name (seeded); centroid draws come from torch.manual_seed(case seed)."""
    import types
    out = {}
    # (name, flavour, n_src, n_tar, fg fraction src, fg fraction tar)
    cases = [("daf_big", "daf", 60, 50, 0.5, 0.5), ("daf_pad", "daf", 30, 12, 0.25, 0.5), ("daf_exact20", "daf", 40, 8, 0.5, 0.5),
             ("daf_src_has_no_fg", "daf", 16, 24, 0.0, 0.5), ("daf_nothing_bg", "daf", 10, 10, 1.0, 1.0),
             ("maf_mixed", "maf", 48, 32, 0.4, 0.6), ("maf_src_has_no_fg", "maf", 20, 20, 0.0, 0.7),
             ("deep_mixed", "deep", 64, 64, 0.5, 0.5), ("deep_src_has_no_bg", "deep", 12, 30, 1.0, 0.3)]
    for idx, (name, flavour, ns_, nt_, fs, ft) in enumerate(cases):
        fns = ref_loader.load_group_loss(flavour)
        cls_head = ns.instance.InstanceAlignmentHead_DAF if flavour == "deep" else ns.instance.InstanceAlignmentHead
        me = types.SimpleNamespace()
        me.local_da_fore = seeded.fill_state_(cls_head().float().eval(), idx, prefix=f"group.{name}.fore.")
        me.local_da_back = seeded.fill_state_(cls_head().float().eval(), idx, prefix=f"group.{name}.back.")
        me.criterion_fl = ns.focal.FocalLoss()
        me.criterion = torch.nn.CrossEntropyLoss()
        me.group = types.MethodType(fns["group"], me)
        me.complete = types.MethodType(fns["complete"], me)
        feats, cls = [], []
        for d, (n, frac) in enumerate(((ns_, fs), (nt_, ft))):
            feats.append(torch.relu(seeded.seeded_tensor(f"group.{name}.feat{d}", (n, 1024), idx)))
            z = seeded.seeded_tensor(f"group.{name}.cls{d}", (n, 2), idx, scale=2.0)
            nfg = int(round(frac * n))
            lo, hi = torch.minimum(z[:, 0], z[:, 1]), torch.maximum(z[:, 0], z[:, 1]) + 0.05
            z = torch.stack([torch.where(torch.arange(n) < nfg, hi, lo), torch.where(torch.arange(n) < nfg, lo, hi)], 1)
            cls.append(z[torch.randperm(n, generator=torch.Generator().manual_seed(idx * 10 + d))])
        if name == "daf_exact20":     # exactly 20 foreground RoIs in the source image
            assert int((torch.softmax(cls[0], -1)[:, 0] >= 0.5).sum()) == 20
        torch.manual_seed(1000 + idx)
        with torch.no_grad():
            val = fns["group_local_da_loss"](me, feats, 0.2, cls)
        # features are not stored: relu(seeded_tensor(f"group.{name}.feat{d}", (n, 1024), seed)), see tests/helpers.group_case
        out[name] = {"flavour": flavour, "seed": idx, "rng_seed": 1000 + idx, "cls": cls, "loss": float(val)}
        print(f"group_local_da_loss {name}: {val:.6f}")
    torch.save(out, os.path.join(OUT, "group_local_da_loss.pt"))


def sampler_cases():
    """Index streams of the reference's BatchSchedulerSampler for seeded epochs (torch.manual_seed(seed) before iter())."""
    from torch.utils.data import ConcatDataset, TensorDataset
    cls = ref_loader.load_batch_sampler()
    out = []
    for sizes, spg, seed in [((7, 4), 2, 0), ((5, 9), 4, 1), ((6, 6), 2, 2), ((3, 10), 6, 3)]:
        ds = ConcatDataset([TensorDataset(torch.zeros(n)) for n in sizes])
        smp = cls(ds, samples_per_gpu=spg)
        torch.manual_seed(seed)
        idx = list(iter(smp))
        out.append({"sizes": list(sizes), "samples_per_gpu": spg, "seed": seed, "len": len(smp), "indices": idx})
        print(f"sampler {sizes} spg={spg}: {len(idx)} indices, len()={len(smp)}")
    import json
    json.dump(out, open(os.path.join(OUT, "batch_scheduler_sampler.json"), "w"))


def state_dict_surface(ns):
    """Key -> shape of the reference DA backbones (trunk + DA heads) and instance heads: the checkpoint surface
    (SURVEY.md Appendix C)."""
    import json
    common = dict(depth=50, num_stages=4, strides=(1, 2, 2, 1), dilations=(1, 1, 1, 2), out_indices=(3,),
                  frozen_stages=1, norm_eval=True, style="pytorch")
    out = {}
    for name, cls in (("ResNet_DAF", ns.daf_org.ResNet_DAF), ("ResNet_DA", ns.maf.ResNet_DA),
                      ("ResNet_DA_CBAM", ns.cbam.ResNet_DA_CBAM), ("ResNet_DA_Deep", ns.deep.ResNet_DA_Deep)):
        out[name] = {k: list(v.shape) for k, v in cls(**common).state_dict().items()}
    out["InstanceAlignmentHead"] = {k: list(v.shape) for k, v in ns.instance.InstanceAlignmentHead().state_dict().items()}
    out["RoILocalAlignmentHead"] = {k: list(v.shape) for k, v in ns.local_da.LocalAlignmentHead(2048).state_dict().items()}
    out["InstanceAlignmentHead_DAF"] = {k: list(v.shape) for k, v in ns.instance.InstanceAlignmentHead_DAF().state_dict().items()}
    json.dump(out, open(os.path.join(OUT, "state_dict_surface.json"), "w"))
    print("state_dict surface:", {k: len(v) for k, v in out.items()})


def main():
    os.makedirs(OUT, exist_ok=True)
    ns = ref_loader.load()
    state_dict_surface(ns)
    roi_align_case()
    fm = seeded.feature_map
    head_case("img_alignment", ns.daf_org.ImgAlignmentHead(64), fm("x.img", (2, 64, 6, 10)))
    head_case("local_alignment", ns.cbam.LocalAlignmentHead(64), fm("x.local", (2, 64, 6, 10)))
    head_case("global_alignment_cbam", ns.cbam.GlobalAlignmentHead(64), fm("x.global", (2, 64, 12, 20)))
    head_case("global_alignment_deep", ns.deep.GlobalAlignmentHead(64), fm("x.globald", (2, 64, 13, 19)))
    head_case("srm", ns.maf.SRM(64), fm("x.srm", (2, 64, 6, 10)))
    head_case("non_local_alignment", ns.deep.NonLocalAlignmentHead(64), fm("x.nla", (2, 64, 4, 6)))
    head_case("instance_alignment", ns.instance.InstanceAlignmentHead(), fm("x.ins", (24, 1024)))
    head_case("roi_local_alignment", ns.local_da.LocalAlignmentHead(64), fm("x.roil", (12, 64, 7, 7)))
    head_case("instance_alignment_daf", ns.instance.InstanceAlignmentHead_DAF(), fm("x.insd", (24, 1024)))
    backbone_loss_cases(ns)
    focal_case(ns)
    bbox_head_loss_case(ns)
    group_loss_cases(ns)
    sampler_cases()
    total = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print(f"golden fixtures: {total / 1e6:.2f} MB in {OUT}")


if __name__ == "__main__":
    main()
