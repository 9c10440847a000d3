"""ORACLE — test infrastructure only; never imported by the product path.

CPU restatement (plain PyTorch, fp32 or fp64, any device torch supports) of the reference's DA
hot path.  Every function cites the reference lines it follows (paths under /root/reference).
The head functions take the head's state_dict (identical keys to the reference classes), so the
same seeded parameters drive the reference classes, this oracle and the CUDA modules.

Parity status: PINNED — tests/test_oracle_cpu.py checks every function here against golden
vectors produced by the reference's own classes (oracle/make_golden.py imports them in place
from /root/reference under an mmcv stub) and checks each closed-form loss against a literal
transcription of the reference's Python loops.
"""
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# GRL — mmdet/models/roi_heads/instance_da.py:14-23
# --------------------------------------------------------------------------------------
class _GRL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w):
        ctx.w = w
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return ctx.w * g.clone(), None


def grl(x, weight=-1.0):
    return _GRL.apply(x, weight)


class _STEQuant(torch.autograd.Function):
    """bf16 storage emulation: round in forward, identity in backward (the CUDA bf16 engine stores
    activations / operands in bf16 and treats the rounding as identity in its backward)."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


def _q(x, q):
    """q=None: exact math.  q='bf16': emulate the bf16 engine's storage points."""
    return x if q is None else _STEQuant.apply(x)


def _bn_eval(x, sd, p, eps=1e-5):
    """nn.BatchNorm2d in eval mode (Q9: norm_eval=True keeps DA-head BN frozen while training)."""
    shape = (1, -1) + (1,) * (x.dim() - 2)
    return (x - sd[p + ".running_mean"].view(shape)) * torch.rsqrt(sd[p + ".running_var"].view(shape) + eps) \
        * sd[p + ".weight"].view(shape) + sd[p + ".bias"].view(shape)


def _drop(x, mask):
    """F.dropout(p=0.5, training=True) with an injected keep-mask (RNG streams cannot be matched, Q9)."""
    return x if mask is None else x * mask.to(x.dtype) * 2.0


# --------------------------------------------------------------------------------------
# H1 ImgAlignmentHead — mmdet/models/backbones/resnet_da_daf_org.py:120-133
# --------------------------------------------------------------------------------------
def img_alignment_head(x, sd, q=None, relu=F.relu):
    """`relu`: the activation (default F.relu).  ReLU has no derivative at 0: a test that compares GRADIENTS of a large
    tensor passes a callable that takes the CUDA path's on/off decision for pre-activations within rounding of 0
    (tests/helpers.ReluLikeCuda) -- either subgradient is valid there, and one flipped unit moves whole gradient rows."""
    x = _q(grl(x), q)
    x = _q(relu(F.conv2d(x, _q(sd["conv1.weight"], q), sd["conv1.bias"])), q)
    return relu(F.conv2d(x, sd["conv2.weight"], sd["conv2.bias"]))


# --------------------------------------------------------------------------------------
# H2 LocalAlignmentHead — mmdet/models/backbones/resnet_da_cbam.py:77-115
# --------------------------------------------------------------------------------------
def local_alignment_head(x, sd, masks=(None, None), q=None):
    x = _q(grl(x), q)
    x = _q(_drop(F.relu(_bn_eval(F.conv2d(x, _q(sd["conv1.weight"], q)), sd, "bn1")), masks[0]), q)
    x = _q(_drop(F.relu(_bn_eval(F.conv2d(x, _q(sd["conv2.weight"], q)), sd, "bn2")), masks[1]), q)
    return F.conv2d(x, sd["conv3.weight"])


# --------------------------------------------------------------------------------------
# H3 GlobalAlignmentHead — resnet_da_cbam.py:117-195 (dead Res-CBAM branch of :176-179 skipped,
# Q10: conv4 consumes `res`, so the branch cannot influence the output) and
# resnet_da_deep.py:206-303 (same head without the branch)
# --------------------------------------------------------------------------------------
def global_alignment_head(x, sd, masks=(None, None, None, None), q=None):
    x = _q(grl(x), q)
    res = _q(_drop(F.relu(_bn_eval(F.conv2d(x, _q(sd["conv1.weight"], q), None, 2, 1), sd, "bn1")), masks[0]), q)
    x = _q(_drop(F.relu(_bn_eval(F.conv2d(res, _q(sd["conv4.weight"], q), None, 2, 1), sd, "bn4")), masks[1]), q)
    x = _q(_drop(F.relu(_bn_eval(F.conv2d(x, _q(sd["conv5.weight"], q), None, 2, 1), sd, "bn5")), masks[2]), q)
    x = F.avg_pool2d(x, (x.size(2), x.size(3))).view(x.size(0), -1)
    x = _drop(F.relu(F.linear(x, sd["fc1.weight"], sd["fc1.bias"])), masks[3])
    return F.linear(x, sd["fc2.weight"], sd["fc2.bias"])


# --------------------------------------------------------------------------------------
# I3 RoI-conv LocalAlignmentHead — mmdet/models/roi_heads/local_da.py:47-86 (3x3 s2 x3 on the [k,C,7,7] RoI features,
# BN (eval) + ReLU + dropout, global average pool, FC 512->2, sigmoid)
# --------------------------------------------------------------------------------------
def roi_local_alignment_logits(x, sd, masks=(None, None, None), q=None):
    x = _q(grl(x), q)
    x = _q(_drop(F.relu(_bn_eval(F.conv2d(x, _q(sd["conv1.weight"], q), None, 2, 1), sd, "bn1")), masks[0]), q)
    x = _q(_drop(F.relu(_bn_eval(F.conv2d(x, _q(sd["conv2.weight"], q), None, 2, 1), sd, "bn2")), masks[1]), q)
    x = _q(_drop(F.relu(_bn_eval(F.conv2d(x, _q(sd["conv3.weight"], q), None, 2, 1), sd, "bn3")), masks[2]), q)
    x = F.avg_pool2d(x, (x.size(2), x.size(3))).view(x.size(0), -1)
    return F.linear(x, sd["fc.weight"], sd["fc.bias"])


def roi_local_alignment_head(x, sd, masks=(None, None, None), q=None):
    return torch.sigmoid(roi_local_alignment_logits(x, sd, masks, q))


# --------------------------------------------------------------------------------------
# H4 SRM — mmdet/models/backbones/resnet_da.py:83-104 (padding=1 on the 1x1, padding=3 on the 3x3, Q12)
# --------------------------------------------------------------------------------------
def srm_logits(x, sd, masks=(None, None), q=None):
    x = _q(grl(x), q)
    x = _q(_drop(F.relu(_bn_eval(F.conv2d(x, _q(sd["conv1.weight"], q), sd["conv1.bias"], 1, 1), sd, "bn1")), masks[0]), q)
    x = _q(_drop(F.relu(_bn_eval(F.conv2d(x, _q(sd["conv2.weight"], q), sd["conv2.bias"], 1, 3), sd, "bn2")), masks[1]), q)
    x = F.avg_pool2d(x, (x.size(2), x.size(3))).view(x.size(0), -1)
    return F.linear(x, sd["fc.weight"], sd["fc.bias"])


def srm(x, sd, masks=(None, None), q=None):
    return torch.sigmoid(srm_logits(x, sd, masks, q))


# --------------------------------------------------------------------------------------
# NonLocalBlock — mmdet/models/roi_heads/instance_da.py:150-192 (softmax over dim=1 of [b,q,k], Q11)
# --------------------------------------------------------------------------------------
def non_local_block(x, sd, p="", q=None):
    b, c, h, w = x.shape
    ic = c // 2
    x_phi = _q(F.conv2d(x, _q(sd[p + "conv_phi.weight"], q)), q).view(b, ic, -1)
    x_theta = _q(F.conv2d(x, _q(sd[p + "conv_theta.weight"], q)), q).view(b, ic, -1).permute(0, 2, 1)
    x_g = _q(F.conv2d(x, _q(sd[p + "conv_g.weight"], q)), q).view(b, ic, -1).permute(0, 2, 1)
    att = _q(torch.softmax(torch.matmul(x_theta, x_phi), dim=1), q)
    y = _q(torch.matmul(att, x_g), q).permute(0, 2, 1).contiguous().view(b, ic, h, w)
    # q='bf16': the residual SUM is the stored bf16 activation on the CUDA path (the fused epilogue adds x to the fp32
    # accumulator and rounds once; it feeds the next GEMM as an operand)
    return _q(F.conv2d(y, _q(sd[p + "conv_mask.weight"], q)) + x, q)


# H5 NonLocalAlignmentHead — mmdet/models/backbones/resnet_da_deep.py:122-164
def non_local_alignment_head(x, sd, mask=None, q=None):
    x = _q(grl(x), q)
    x = _q(_drop(F.relu(_bn_eval(F.conv2d(x, _q(sd["conv1.weight"], q)), sd, "bn1")), mask), q)
    return non_local_block(x, sd, "nlb1.", q)


# --------------------------------------------------------------------------------------
# I1 InstanceAlignmentHead — instance_da.py:42-86 ; I2 InstanceAlignmentHead_DAF — :103-131
# --------------------------------------------------------------------------------------
def instance_alignment_logits(x, sd, masks=(None, None), q=None, relu=F.relu):
    x = _q(grl(x), q)
    x = x.unsqueeze(0).permute(0, 2, 1).contiguous().unsqueeze(2)  # [1,C,1,k]
    x = non_local_block(x, sd, "nlb.", q)
    x = x.permute(3, 1, 0, 2).contiguous().squeeze(-1).squeeze(-1)  # [k,C]
    x = _q(_drop(relu(F.linear(x, _q(sd["fc1.weight"], q), sd["fc1.bias"])), masks[0]), q)
    x = _q(_drop(relu(F.linear(x, _q(sd["fc2.weight"], q), sd["fc2.bias"])), masks[1]), q)
    return F.linear(x, _q(sd["fc3.weight"], q), sd["fc3.bias"])


def instance_alignment_daf_logits(x, sd, masks=(None, None), q=None):
    x = _q(grl(x), q)
    x = _q(_drop(F.relu(F.linear(x, _q(sd["fc1.weight"], q), sd["fc1.bias"])), masks[0]), q)
    x = _q(_drop(F.relu(F.linear(x, _q(sd["fc2.weight"], q), sd["fc2.bias"])), masks[1]), q)
    return F.linear(x, _q(sd["fc3.weight"], q), sd["fc3.bias"])


# --------------------------------------------------------------------------------------
# losses — closed forms and literal loop transcriptions
# --------------------------------------------------------------------------------------
def daf_image_loss_loop(patch_feat, gt_domain):
    """Literal: resnet_da_daf_org.py:816-822 (note: the loop variable only selects the branch)."""
    losses = []
    for i in range(len(patch_feat)):
        if gt_domain[i] == 0:
            losses.append(0.5 * torch.mean(torch.sigmoid(patch_feat) ** 2))
        elif gt_domain[i] == 1:
            losses.append(0.5 * torch.mean(torch.sigmoid(1 - patch_feat) ** 2))
    return sum(losses)


def daf_image_loss(patch_feat, gt_domain):
    """L1 closed form."""
    n_src = int((gt_domain == 0).sum())
    n_tgt = int((gt_domain == 1).sum())
    return 0.5 * (n_src * torch.mean(torch.sigmoid(patch_feat) ** 2) + n_tgt * torch.mean(torch.sigmoid(1 - patch_feat) ** 2))


def patch_loss_loop(local_feat, gt_domain):
    """Literal: resnet_da_cbam.py:971-979."""
    losses = []
    for i, patch_feat in enumerate(local_feat):
        if gt_domain[i] == 0:
            losses.append(0.5 * torch.mean(torch.sigmoid(patch_feat) ** 2))
        elif gt_domain[i] == 1:
            losses.append(0.5 * torch.mean(torch.sigmoid(1 - patch_feat) ** 2))
    return sum(losses)


def patch_loss(local_feat, gt_domain):
    """L2 closed form."""
    n = local_feat.shape[0]
    flat = local_feat.reshape(n, -1)
    d = gt_domain.view(n, 1).to(flat.dtype)
    src = 0.5 * torch.mean(torch.sigmoid(flat) ** 2, dim=1)
    tgt = 0.5 * torch.mean(torch.sigmoid(1 - flat) ** 2, dim=1)
    return (torch.where(d.view(-1) == 0, src, torch.zeros_like(src)) + torch.where(d.view(-1) == 1, tgt, torch.zeros_like(tgt))).sum()


def ce2(u, labels):
    """L3 / L4: nn.CrossEntropyLoss()(u, labels) (resnet_da_cbam.py:966-968; on sigmoid outputs at
    resnet_da.py:846-848 and DAFaster_rcnn_Orig.py:185-186 -> pass u = sigmoid(z))."""
    return F.cross_entropy(u, labels.long())


def focal2(u, labels, gamma=2.0, alpha=0.25):
    """L6: py_sigmoid_focal_loss, mmdet/models/losses/focal_loss.py:12-57 with the one-hot target of
    FocalLoss.forward (:166-168); mean over k*2."""
    p = u.sigmoid()
    t = F.one_hot(labels.long(), num_classes=3)[:, :2].type_as(u)
    pt = (1 - p) * t + p * (1 - t)
    fw = (alpha * t + (1 - alpha) * (1 - t)) * pt.pow(gamma)
    return (F.binary_cross_entropy_with_logits(u, t, reduction="none") * fw).mean()


def consistency_loss_loop(imgs_feat, ins_preds, ins_labels):
    """Literal: DAFaster_rcnn_Orig.py:161-175 (device-parametrised)."""
    loss = torch.zeros((), dtype=imgs_feat.dtype, device=imgs_feat.device)
    for i, img_feat in enumerate(imgs_feat):
        img_logit = torch.sigmoid(imgs_feat)
        I = torch.nonzero(img_logit).size()[0]
        img_logit = img_logit.sum() / I
        for j, ins_label in enumerate(ins_labels):
            if ins_label == i:
                ins_logit = torch.sigmoid(ins_preds[j])
                loss = loss + torch.dist(img_logit, ins_logit[i], p=2)
    return loss


def consistency_loss(imgs_feat, ins_preds, ins_labels):
    """L7 closed form: sum_r |mean(sigmoid(imgs_feat)) - sigmoid(ins_preds[r, label_r])|."""
    m = torch.sigmoid(imgs_feat).mean()
    lab = ins_labels.long()
    valid = (lab >= 0) & (lab < imgs_feat.shape[0]) & (lab < 2)
    s = torch.sigmoid(ins_preds.gather(1, lab.clamp(0, 1).view(-1, 1)).view(-1))
    return (torch.abs(m - s) * valid.to(s.dtype)).sum()


# --------------------------------------------------------------------------------------
# RoI construction — mmdet/core/bbox/transforms.py:59-78
# --------------------------------------------------------------------------------------
def bbox2roi(bbox_list):
    rois_list = []
    for img_id, bboxes in enumerate(bbox_list):
        if bboxes.size(0) > 0:
            img_inds = bboxes.new_full((bboxes.size(0), 1), img_id)
            rois = torch.cat([img_inds, bboxes[:, :4]], dim=-1)
        else:
            rois = bboxes.new_zeros((0, 5))
        rois_list.append(rois)
    return torch.cat(rois_list, 0)


# --------------------------------------------------------------------------------------
# Composite: the DAF-Org DA losses on given features (DAFaster_rcnn_Orig.py:143-157 weights)
# --------------------------------------------------------------------------------------
def daf_org_da_losses(c5, bbox_feats, gt_domain, sd_img, sd_ins, lam=(0.1, 0.1, 0.1)):
    """c5 [N,C,H,W]; bbox_feats = [feat_src [R0,1024], feat_tgt [R1,1024]] ->
    dict(globle_da_loss, local_da_loss, consistency_loss) exactly as forward_train weights them."""
    img_feat = img_alignment_head(c5, sd_img)
    g_loss = daf_image_loss(img_feat, gt_domain)
    labels = torch.cat([torch.zeros(len(bbox_feats[0])), torch.ones(len(bbox_feats[1]))]).long().to(c5.device)
    pred = torch.sigmoid(instance_alignment_logits(torch.cat(bbox_feats, 0), sd_ins))
    l_loss = ce2(pred, labels)
    c_loss = consistency_loss(img_feat, pred, labels)
    return dict(globle_da_loss=lam[0] * g_loss, local_da_loss=lam[1] * l_loss, consistency_loss=lam[2] * c_loss), \
        img_feat, pred


# --------------------------------------------------------------------------------------
# L5 group_local_da_loss — mmdet/models/detectors/DAFaster_rcnn.py:198-327 (DAF),
# MAFaster_rcnn.py:204-299 (MAF), DAFaster_rcnn_Deep.py:232-329 (Deep).  Pinned to the reference's own
# methods run on CPU (oracle/ref_loader.load_group_loss, golden group_local_da_loss.pt).
# What the reference code actually computes (each point is visible in the cited lines):
#   * RoI i is "foreground" iff softmax(bbox_cls[d][i])[0] >= 0.5                          (:240, :267)
#   * DAF only: a group with more than k=20 members is replaced by the k-means "centroids" (:219-223) -
#     but cluster.forward never writes the updated centroids back (`local = ...` rebinds a loop variable,
#     models/utils/cluster.py:138-140), so the result is its random initialisation: 10 draws of randn(1024)
#     (cluster.py:95-98), in the order fg_src, bg_src, fg_tar, bg_tar; a group of exactly 20 is kept; a smaller
#     one is padded to 20 with copies of the member whose score has the largest softmax over the group (:198-210)
#   * `len(fg_src)!=0 & len(fg_tar)!=0` parses as `len(fg_src) != (0 & len(fg_tar)) != 0`, which is always
#     False: the source group is used alone whenever it is non-empty (labels 0), else the target group
#     (labels 1)                                                                            (:283-305)
#   * DAF/MAF call the head once per row with a [1,1024] input (:311-313): the NonLocalBlock then attends over a
#     single token, i.e. y = x + W_mask W_g x; Deep passes the whole batch to the FC-only head (:316-317)
#   * DAF: FocalLoss(gamma 2, alpha .25, mean) on the sigmoid outputs; MAF/Deep: CrossEntropy on them
#   * the sum of the two losses is returned through .item(): a Python float, no gradient                (:325)
# --------------------------------------------------------------------------------------
def instance_alignment_single_token_logits(x, sd, q=None):
    """InstanceAlignmentHead applied row by row ([1,1024] inputs): attention over one token is the identity."""
    x = _q(grl(x), q)
    g = _q(F.linear(x, _q(sd["nlb.conv_g.weight"].flatten(1), q)), q)
    x = _q(F.linear(g, _q(sd["nlb.conv_mask.weight"].flatten(1), q)), q) + x
    x = _q(F.relu(F.linear(x, _q(sd["fc1.weight"], q), sd["fc1.bias"])), q)
    x = _q(F.relu(F.linear(x, _q(sd["fc2.weight"], q), sd["fc2.bias"])), q)
    return F.linear(x, _q(sd["fc3.weight"], q), sd["fc3.bias"])


def draw_centroids(dim=1024, n=10):
    """The reference's centroid initialisation (cluster.py:95-98): n separate torch.randn([dim]) draws from the global RNG."""
    return torch.stack([torch.randn([dim]) for _ in range(n)], 0)


def group_features(feats, scores, k=20, draw=draw_centroids):
    n = feats.shape[0]
    if n > k:
        return draw(feats.shape[1]).to(feats)
    if n == k:
        return feats
    top = torch.argmax(torch.softmax(scores, dim=-1), dim=0)
    return torch.cat([feats, feats[top].unsqueeze(0).expand(k - n, -1)], 0)


def group_local_da_loss(bbox_feats, bbox_cls, sd_fore, sd_back, flavour="daf", k=20, draw=draw_centroids, q=None):
    groups = {}
    for d in (0, 1):                                     # source first, then target: the order of the RNG draws
        p = torch.softmax(bbox_cls[d], dim=-1)
        fg = p[:, 0] >= 0.5
        for name, m, score in (("fg", fg, p[:, 0]), ("bg", ~fg, p[:, 1])):
            f = bbox_feats[d][m]
            if f.shape[0] and flavour == "daf":
                f = group_features(f, score[m], k, draw)
            groups[(name, d)] = f
    total = bbox_feats[0].new_zeros(())
    for name, sd in (("fg", sd_fore), ("bg", sd_back)):
        src, tar = groups[(name, 0)], groups[(name, 1)]
        if src.shape[0]:
            f, label = src, 0
        elif tar.shape[0]:
            f, label = tar, 1
        else:
            continue
        z = instance_alignment_daf_logits(f, sd, q=q) if flavour == "deep" else instance_alignment_single_token_logits(f, sd, q=q)
        pred = torch.sigmoid(z)
        labels = torch.full((f.shape[0],), label, dtype=torch.long)
        total = total + (focal2(pred, labels) if flavour == "daf" else ce2(pred, labels))
    return total.detach()


# --------------------------------------------------------------------------------------
# R3 StandardRoIHeadDA_v5 — mmdet/models/roi_heads/standard_roi_head_da_v5.py:162-227 and the box-head loss it calls
# (bbox_heads/bbox_head.py:256-315 with the DA configs' CrossEntropyLoss(use_sigmoid=True) + SmoothL1Loss(beta=1),
# faster_rcnn_r50_torch_daf.py:56-58).  Pinned to the reference's own loss classes by tests/golden/bbox_head_loss.pt.
# --------------------------------------------------------------------------------------
def bbox_head_loss(cls_score, bbox_pred, labels, targets, pos_mask, num_classes, beta=1.0, reg_class_agnostic=False):
    """loss_cls = sum BCEWithLogits(cls_score, onehot_{C+1}(labels)) / R   (weight_reduce_loss with avg_factor = #RoIs,
    losses/cross_entropy_loss.py:100-114);  loss_bbox = sum SmoothL1(pred[pos, label], target[pos]) / R
    (avg_factor = bbox_targets.size(0), bbox_head.py:309);  acc = top-1 in percent (losses/accuracy.py)."""
    R = max(cls_score.shape[0], 1)
    onehot = F.one_hot(labels, num_classes + 1).to(cls_score.dtype)
    out = {"loss_cls": F.binary_cross_entropy_with_logits(cls_score, onehot, reduction="sum") / R,
           "acc": (cls_score.argmax(1) == labels).to(cls_score.dtype).mean() * 100}
    if bool(pos_mask.any()):
        pred = bbox_pred[pos_mask] if reg_class_agnostic else bbox_pred.view(bbox_pred.shape[0], -1, 4)[pos_mask, labels[pos_mask]]
        d = (pred - targets[pos_mask]).abs()
        out["loss_bbox"] = torch.where(d < beta, 0.5 * d * d / beta, d - 0.5 * beta).sum() / R
    else:
        out["loss_bbox"] = bbox_pred.sum() * 0
    return out


def roi_head_da_v5(pooled_per_img, sampled, sd, num_classes, gt_da, q=None):
    """What _bbox_forward_train returns for a (source, target) pair: per image i the RoI features pooled from image i
    (`pooled_per_img[i]` [n_i, C, 7, 7]; the RoIAlign itself is oracle/roi_align_ref.c) go through forward_train_da
    (convfc_bbox_head.py:198-237: flatten, FC+ReLU x2, fc_cls, fc_reg); the box loss is taken on the SOURCE image only
    (:206-211).  sampled[i] = (boxes, labels, targets, pos_mask) of image i; sd = state_dict of the bbox head.
    -> (losses, bbox_feats=[feat_src, feat_tar], bbox_cls=[cls_src, cls_tar])."""
    feats, clss, regs = [], [], []
    for x in pooled_per_img:
        t = _q(x.flatten(1), q)
        t = _q(F.relu(F.linear(t, _q(sd["shared_fcs.0.weight"], q), sd["shared_fcs.0.bias"])), q)
        t = _q(F.relu(F.linear(t, _q(sd["shared_fcs.1.weight"], q), sd["shared_fcs.1.bias"])), q)
        feats.append(t)
        clss.append(F.linear(t, sd["fc_cls.weight"], sd["fc_cls.bias"]))
        regs.append(F.linear(t, sd["fc_reg.weight"], sd["fc_reg.bias"]))
    losses = {}
    src = [i for i, d in enumerate(gt_da) if int(d) == 0]
    if src:
        i = src[0]
        losses = bbox_head_loss(clss[i], regs[i], sampled[i][1], sampled[i][2], sampled[i][3], num_classes)
    return losses, feats, clss
