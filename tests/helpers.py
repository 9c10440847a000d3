"""Shared helpers of the parity tests."""
import torch

import unsupervised_domain_adaptation_object_detection_implementation_b200 as uda
from unsupervised_domain_adaptation_object_detection_implementation_b200 import da_heads
from oracle import da_oracle, seeded

# golden name -> (module factory of THIS repo, oracle function(x, sd), input is 4-D feature map?)
HEADS = {
    "img_alignment": (lambda: da_heads.ImgAlignmentHead(64), da_oracle.img_alignment_head),
    "local_alignment": (lambda: da_heads.LocalAlignmentHead(64), da_oracle.local_alignment_head),
    "global_alignment_cbam": (lambda: da_heads.GlobalAlignmentHead(64), da_oracle.global_alignment_head),
    "global_alignment_deep": (lambda: da_heads.GlobalAlignmentHeadDeep(64), da_oracle.global_alignment_head),
    "srm": (lambda: da_heads.SRM(64), da_oracle.srm),
    "non_local_alignment": (lambda: da_heads.NonLocalAlignmentHead(64), da_oracle.non_local_alignment_head),
    "instance_alignment": (lambda: da_heads.InstanceAlignmentHead(),
                           lambda x, sd, q=None: torch.sigmoid(da_oracle.instance_alignment_logits(x, sd, q=q))),
    "roi_local_alignment": (lambda: da_heads.RoILocalAlignmentHead(64), da_oracle.roi_local_alignment_head),
    "instance_alignment_daf": (lambda: da_heads.InstanceAlignmentHead_DAF(),
                               lambda x, sd, q=None: torch.sigmoid(da_oracle.instance_alignment_daf_logits(x, sd, q=q))),
}


def build_head(name, seed=0):
    m = HEADS[name][0]().float().eval()
    seeded.fill_state_(m, seed, prefix=name + ".")
    return m


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def check_summary(grad, summ, tol):
    """Compare a gradient tensor against the (possibly subsampled) golden record."""
    g = grad.detach().float().cpu().reshape(-1)
    if "full" in summ:
        ref = summ["full"]
        scale = ref.abs().max().clamp_min(1e-30)
        return float((g - ref).abs().max() / scale)
    ref = summ["sample"]
    scale = max(float(ref.abs().max()), 1e-30)
    e1 = float((g[summ["idx"]] - ref).abs().max()) / scale
    e2 = abs(float(g.norm()) - summ["norm"]) / max(summ["norm"], 1e-30)
    return max(e1, e2)


def group_case(rec, name):
    """Inputs of one golden case of group_local_da_loss.pt: features are regenerated from their seeds."""
    feats = [torch.relu(seeded.seeded_tensor(f"group.{name}.feat{d}", (rec["cls"][d].shape[0], 1024), rec["seed"])) for d in (0, 1)]
    return feats, rec["cls"]


def group_heads(rec, name):
    cls_head = da_heads.InstanceAlignmentHead_DAF if rec["flavour"] == "deep" else da_heads.InstanceAlignmentHead
    fore = seeded.fill_state_(cls_head().float().eval(), rec["seed"], prefix=f"group.{name}.fore.")
    back = seeded.fill_state_(cls_head().float().eval(), rec["seed"], prefix=f"group.{name}.back.")
    return fore, back


class ReluLikeCuda:
    """ReLU for the ORACLE side of a gradient comparison at large sizes.  The function is evaluated exactly; only where a
    pre-activation lies within `delta` (relative to the tensor's max magnitude) of 0 -- where the two sides' rounding decides
    whether the unit is on, and either subgradient of ReLU is valid -- the oracle takes the on/off decision the CUDA path
    made (`cuda_acts`: the CUDA path's post-activation tensors, in call order, already in the oracle's layout).
    `flips` counts the near-zero units where that changed the mask; `outside` counts disagreements OUTSIDE the delta band
    (must be 0: that would be a real error, and the test asserts it)."""

    def __init__(self, cuda_acts, delta):
        self.acts, self.delta, self.i, self.flips, self.outside, self.near = list(cuda_acts), delta, 0, 0, 0, 0

    def __call__(self, z):
        on_cuda = (self.acts[self.i].to(z.device) > 0)
        self.i += 1
        assert on_cuda.shape == z.shape, (on_cuda.shape, z.shape)
        zd = z.detach()
        near = zd.abs() <= self.delta * zd.abs().max()
        on_self = zd > 0
        self.near += int(near.sum())
        self.flips += int((near & (on_cuda != on_self)).sum())
        self.outside += int((~near & (on_cuda != on_self)).sum())
        return z * torch.where(near, on_cuda, on_self).to(z.dtype)
