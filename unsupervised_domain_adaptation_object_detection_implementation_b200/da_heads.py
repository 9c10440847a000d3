"""Domain-classifier heads of the reference, re-hosted on libda_b200 kernels.

Class names, constructor arguments, parameter names (state_dict keys) and return values mirror
the reference modules so that `da_configs` and reference checkpoints keep working:

  GradientScalarLayer        mmdet/models/roi_heads/instance_da.py:14-40
  ImgAlignmentHead           mmdet/models/backbones/resnet_da_daf_org.py:120-146   (H1)
  LocalAlignmentHead         mmdet/models/backbones/resnet_da_cbam.py:77-115       (H2)
  GlobalAlignmentHead        mmdet/models/backbones/resnet_da_cbam.py:117-214      (H3, CBAM flavour)
  GlobalAlignmentHeadDeep    mmdet/models/backbones/resnet_da_deep.py:206-303      (H3, Deep flavour)
  SRM                        mmdet/models/backbones/resnet_da.py:83-118            (H4)
  NonLocalBlock              mmdet/models/roi_heads/instance_da.py:150-192         (Q11)
  NonLocalAlignmentHead      mmdet/models/backbones/resnet_da_deep.py:122-164      (H5)
  InstanceAlignmentHead      mmdet/models/roi_heads/instance_da.py:42-101          (I1)
  InstanceAlignmentHead_DAF  mmdet/models/roi_heads/instance_da.py:103-148         (I2)

The nn.Conv2d / nn.Linear / nn.BatchNorm2d children are PARAMETER CONTAINERS only (identical
keys and shapes); the arithmetic runs in the fused kernels:
  GRL          -> folded into the first layer's data-gradient epilogue (out_scale = weight)
  conv+BN+ReLU+Dropout -> one implicit-GEMM with fused epilogue (eval-mode BN folded, Q9)
  1-channel terminal conv -> pixel_head (GEMV)
Dropout uses a stateless counter hash (seed drawn from torch's generator per call), so the
RNG stream differs from torch's Philox; `da_dropout_mask` exports the mask.
"""
import torch
import torch.nn as nn

from . import functional as F_


class GradientScalarLayer(nn.Module):
    """instance_da.py:26-40.  Generic (unfused) form; the heads below fold it instead."""

    def __init__(self, weight):
        super().__init__()
        self.weight = weight

    def forward(self, input):
        return F_.gradient_scalar(input, self.weight)

    def __repr__(self):
        return f"{self.__class__.__name__}(weight={self.weight})"


def normal_init(m, mean, stddev):
    m.weight.data.normal_(mean, stddev)


def _channels_last_(conv):
    """Store a conv weight as OHWI physically (logical shape/state_dict key unchanged)."""
    if conv.weight.dim() == 4:
        conv.weight.data = conv.weight.data.contiguous(memory_format=torch.channels_last)
    return conv


def _draw_seed(training, p):
    if not training or p <= 0.0:
        return 0
    return int(torch.randint(0, 2 ** 62, (1,)).item())


def bn_affine(bn, conv_bias=None):
    """Eval-mode BatchNorm as per-channel (scale, shift), with an optional preceding conv bias
    folded in.  Differentiable w.r.t. bn.weight / bn.bias / conv_bias (they keep training in the
    reference even though the statistics are frozen, Q9)."""
    if bn.training:
        raise NotImplementedError(
            "DA-head BatchNorm runs in eval mode in the reference (norm_eval=True + train() override, "
            "resnet_da_cbam.py:995-1004); call .eval() on the BatchNorm or use the backbone's train().")
    inv = torch.rsqrt(bn.running_var + bn.eps)
    scale = bn.weight * inv if bn.affine else inv
    shift = -bn.running_mean * scale
    if bn.affine:
        shift = shift + bn.bias
    if conv_bias is not None:
        shift = shift + conv_bias * scale
    return scale, shift


class _HeadBase(nn.Module):
    grl_weight = -1.0
    drop_p = 0.5

    def _p(self):
        return self.drop_p if self.training else 0.0

    def _layer(self, x, conv, bn=None, relu=True, drop=True, grl=1.0, stride=None, pad=None, out_dtype=None, engine=None):
        if bn is not None:
            scale, shift = bn_affine(bn, conv.bias)
        else:
            scale, shift = None, conv.bias
        p = self._p() if drop else 0.0
        if isinstance(conv, nn.Conv2d):
            stride = conv.stride[0] if stride is None else stride
            pad = conv.padding[0] if pad is None else pad
        else:
            stride, pad = 1, 0
        return F_.dense_layer(x, conv.weight, scale, shift, stride=stride, pad=pad, relu=relu, drop_p=p,
                              seed=_draw_seed(self.training, p), grl=grl, out_dtype=out_dtype, engine=engine)

    def train(self, mode=True):
        super().train(mode)
        # the reference keeps every DA-head BatchNorm in eval mode while training (Q9)
        for m in self.modules():
            if isinstance(m, nn.modules.batchnorm._BatchNorm):
                m.eval()
        return self


def _tiny_engine():
    # [N,C] FC tails (N = 2 images) are latency bound: CUDA-core kernel, fp32
    return "simt_f32"


# --------------------------------------------------------------------------------------
# H1  ImgAlignmentHead (DAF-Org image-level head on C5)
# --------------------------------------------------------------------------------------
class ImgAlignmentHead(_HeadBase):
    def __init__(self, in_channel):
        super().__init__()
        self.grl = GradientScalarLayer(-1.0)
        self.conv1 = _channels_last_(nn.Conv2d(in_channel, 512, kernel_size=1, stride=1, padding=0))
        self.conv2 = nn.Conv2d(512, 1, kernel_size=1, stride=1, padding=0)

    def forward(self, x):
        """x [N,C,H,W] -> img_feat [N,1,H,W] = relu(conv2(relu(conv1(grl(x)))))."""
        a = F_.to_nhwc(x, F_.act_dtype())
        h = self._layer(a, self.conv1, relu=True, drop=False, grl=self.grl.weight)
        logits = F_.pixel_head(h, self.conv2.weight, self.conv2.bias, relu=True)  # [N,H,W]
        return logits.unsqueeze(1)

    def forward_loss(self, x, gt_domain, mode="daf_sq_batch", gamma=2.0, alpha=0.25):
        """(loss, img_feat [N,1,H,W]): the head AND its per-pixel domain loss (L1 by default: resnet_da_daf_org.py:816-822; 'bce' /
        'focal' are the plain per-pixel modes) as conv1 (tcgen05 GEMM, GRL folded into its data gradient) + ONE tail kernel
        (conv2 + bias + ReLU + loss + mean); backward = one tail kernel (loss', conv2 gradients, ReLU mask of conv1) + the
        weight / data gradient GEMMs (functional.conv_pixel_loss -> da_grl_conv_loss_forward/backward)."""
        if not USE_FUSED_TAIL:
            feat = self.forward(x)
            return F_.pixel_domain_loss(feat, gt_domain, True), feat
        a = F_.to_nhwc(x, F_.act_dtype())
        loss, logits = F_.conv_pixel_loss(a, self.conv1.weight, None, self.conv1.bias, self.conv2.weight, self.conv2.bias, gt_domain,
                                          relu=True, grl=self.grl.weight, tail_relu=True, mode=mode, gamma=gamma, alpha=alpha)
        return loss, logits.unsqueeze(1)

    def _init_weights(self):
        normal_init(self.conv1, 0, 0.001)
        normal_init(self.conv2, 0, 0.001)


# --------------------------------------------------------------------------------------
# H2  LocalAlignmentHead (pixel-level head on C3)
# --------------------------------------------------------------------------------------
class LocalAlignmentHead(_HeadBase):
    def __init__(self, in_channels, context=False, grl=True):
        super().__init__()
        self.grl_flag = grl
        self.grl = GradientScalarLayer(-1.0)
        self.drop = nn.Dropout(p=0.5)
        self.conv1 = _channels_last_(nn.Conv2d(in_channels, in_channels, 1, bias=False))
        self.bn1 = nn.BatchNorm2d(in_channels)
        self.conv2 = _channels_last_(nn.Conv2d(in_channels, in_channels, 1, bias=False))
        self.bn2 = nn.BatchNorm2d(in_channels)
        self.conv3 = nn.Conv2d(in_channels, 1, 1, bias=False)
        self.context = context
        self._init_weights()

    def _init_weights(self):
        normal_init(self.conv1, 0, 0.01)
        normal_init(self.conv2, 0, 0.01)
        normal_init(self.conv3, 0, 0.01)

    def forward_loss(self, x, gt_domain, mode="daf_sq_image", gamma=2.0, alpha=0.25):
        """(loss, local_feat [N,1,H,W]): the head AND its per-pixel loss (L2 by default: resnet_da_cbam.py:971-979).  conv1 is
        a plain fused layer; conv2 (+BN+ReLU+dropout) -> conv3 -> loss -> mean run as GEMM + one tail kernel."""
        a = F_.to_nhwc(x, F_.act_dtype())
        h = self._layer(a, self.conv1, self.bn1, grl=self.grl.weight)
        scale, shift = bn_affine(self.bn2, self.conv2.bias)
        p = self._p()
        loss, logits = F_.conv_pixel_loss(h, self.conv2.weight, scale, shift, self.conv3.weight, None, gt_domain, relu=True, drop_p=p,
                                          seed=_draw_seed(self.training, p), tail_relu=False, mode=mode, gamma=gamma, alpha=alpha)
        return loss, logits.unsqueeze(1)

    def forward(self, x):
        a = F_.to_nhwc(x, F_.act_dtype())
        h = self._layer(a, self.conv1, self.bn1, grl=self.grl.weight)
        h = self._layer(h, self.conv2, self.bn2)
        logits = F_.pixel_head(h, self.conv3.weight, None, relu=False)
        if self.context:
            feat = F_.global_avgpool(h).view(h.shape[0], -1, 1, 1)
            return logits.unsqueeze(1), feat
        return logits.unsqueeze(1)


# --------------------------------------------------------------------------------------
# H3  GlobalAlignmentHead
# --------------------------------------------------------------------------------------
class CBAMLayer(nn.Module):
    """resnet_da_cbam.py:227-268.  Parameter container: the Res-CBAM branch of GlobalAlignmentHead
    is dead code in the reference (its result is discarded, Q10), so it is never evaluated."""

    def __init__(self, channel, reduction=16, spatial_kernel=7):
        super().__init__()
        self.mlp = nn.Sequential(nn.Conv2d(channel, channel // reduction, 1, bias=False), nn.ReLU(inplace=True),
                                 nn.Conv2d(channel // reduction, channel, 1, bias=False))
        self.conv = nn.Conv2d(2, 1, kernel_size=spatial_kernel, padding=spatial_kernel // 2, bias=False)


def conv3x3(in_planes, out_planes, stride=1):
    return _channels_last_(nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=1, bias=False))


class GlobalAlignmentHead(_HeadBase):
    """CBAM flavour.  The dead branch (conv2, conv3, CBAM) keeps its parameters for checkpoint
    compatibility; they never receive gradients (static unused-parameter list for DDP)."""

    dead_branch = True

    def __init__(self, in_channel, context=False, grl=True):
        super().__init__()
        self.grl_flag = grl
        self.output_channel = int(in_channel / 4)
        half = int(in_channel / 2)
        self.grl = GradientScalarLayer(-1.0)
        self.conv1 = conv3x3(in_channel, half, stride=2)
        self.bn1 = nn.BatchNorm2d(half)
        if self.dead_branch:
            self.conv2 = nn.Conv2d(half, half, kernel_size=3, stride=1, padding=1)
            self.bn2 = nn.BatchNorm2d(half)
            self.conv3 = nn.Conv2d(half, half, kernel_size=3, stride=1, padding=1)
            self.bn3 = nn.BatchNorm2d(half)
            self.CBAM = CBAMLayer(channel=half)
        self.conv4 = conv3x3(half, self.output_channel, stride=2)
        self.bn4 = nn.BatchNorm2d(self.output_channel)
        self.conv5 = conv3x3(self.output_channel, self.output_channel, stride=2)
        self.bn5 = nn.BatchNorm2d(self.output_channel)
        self.fc1 = nn.Linear(self.output_channel, int(self.output_channel / 2))
        self.fc2 = nn.Linear(int(self.output_channel / 2), 2)
        self.context = context

    def unused_parameters(self):
        if not self.dead_branch:
            return []
        mods = [self.conv2, self.bn2, self.conv3, self.bn3, self.CBAM]
        return [p for m in mods for p in m.parameters()]

    def forward(self, x):
        """x [N,C,H,W] -> raw domain logits [N,2]."""
        a = F_.to_nhwc(x, F_.act_dtype())
        res = self._layer(a, self.conv1, self.bn1, grl=self.grl.weight)
        h = self._layer(res, self.conv4, self.bn4)
        h = self._layer(h, self.conv5, self.bn5)
        feat = F_.global_avgpool(h)  # [N, C/4] fp32
        t = self._layer(feat.view(feat.shape[0], 1, 1, -1), self.fc1, relu=True, drop=True, engine=_tiny_engine())
        z = self._layer(t, self.fc2, relu=False, drop=False, engine=_tiny_engine())
        z = z.view(z.shape[0], -1)
        if self.context:
            return z, feat
        return z

    def _init_weights(self):
        names = ["conv1", "conv4", "conv5", "fc1", "fc2"] + (["conv2", "conv3"] if self.dead_branch else [])
        for n in names:
            normal_init(getattr(self, n), 0, 0.01)


class GlobalAlignmentHeadDeep(GlobalAlignmentHead):
    """resnet_da_deep.py:206-303: same head without the dead Res-CBAM branch."""

    dead_branch = False


# --------------------------------------------------------------------------------------
# I3  RoI-conv LocalAlignmentHead (roi_heads/local_da.py:47-101)
# --------------------------------------------------------------------------------------
class RoILocalAlignmentHead(_HeadBase):
    """The reference's `roi_heads.local_da.LocalAlignmentHead`: GRL -> three 3x3 stride-2 convs (C -> 1024 -> 512 ->
    512) with BN (eval) + ReLU + dropout on the [k,C,7,7] RoI features -> global average pool -> FC 512->2 ->
    sigmoid.  Same state_dict keys; named apart from the backbone `LocalAlignmentHead` it shares a class name with."""

    def __init__(self, in_channel, context=False):
        super().__init__()
        self.output_channel = 512
        self.grl = GradientScalarLayer(weight=-1.0)
        self.conv1 = conv3x3(in_channel, 1024, stride=2)
        self.bn1 = nn.BatchNorm2d(1024)
        self.conv2 = conv3x3(1024, self.output_channel, stride=2)
        self.bn2 = nn.BatchNorm2d(self.output_channel)
        self.conv3 = conv3x3(self.output_channel, self.output_channel, stride=2)
        self.bn3 = nn.BatchNorm2d(self.output_channel)
        self.fc = nn.Linear(self.output_channel, 2)
        self.context = context

    def forward_logits(self, x):
        a = F_.to_nhwc(x, F_.act_dtype())
        h = self._layer(a, self.conv1, self.bn1, grl=self.grl.weight)
        h = self._layer(h, self.conv2, self.bn2)
        h = self._layer(h, self.conv3, self.bn3)
        feat = F_.global_avgpool(h)
        z = self._layer(feat.view(feat.shape[0], 1, 1, -1), self.fc, relu=False, drop=False, engine=_tiny_engine(),
                        out_dtype=torch.float32)
        return z.view(z.shape[0], 2).float()

    def forward(self, x):
        return torch.sigmoid(self.forward_logits(x))

    def _init_weights(self):
        for n in ("conv1", "conv2", "conv3"):
            normal_init(getattr(self, n), 0, 1)


# --------------------------------------------------------------------------------------
# H4  SRM (MAF head)
# --------------------------------------------------------------------------------------
class SRM(_HeadBase):
    def __init__(self, in_channel):
        super().__init__()
        q = int(in_channel / 4)
        self.output_channel = q * 3 * 3
        self.grl = GradientScalarLayer(-1.0)
        # padding=1 on a 1x1 conv and padding=3 on a 3x3 conv are the reference's (Q12)
        self.conv1 = _channels_last_(nn.Conv2d(in_channel, q, kernel_size=1, padding=1, stride=1))
        self.bn1 = nn.BatchNorm2d(q)
        self.conv2 = _channels_last_(nn.Conv2d(q, self.output_channel, kernel_size=3, padding=3))
        self.bn2 = nn.BatchNorm2d(self.output_channel)
        self.fc = nn.Linear(self.output_channel, 2)

    def forward_logits(self, x):
        a = F_.to_nhwc(x, F_.act_dtype())
        h = self._layer(a, self.conv1, self.bn1, grl=self.grl.weight)
        h = self._layer(h, self.conv2, self.bn2)
        feat = F_.global_avgpool(h)
        z = self._layer(feat.view(feat.shape[0], 1, 1, -1), self.fc, relu=False, drop=False, engine=_tiny_engine())
        return z.view(z.shape[0], -1)

    def forward(self, x):
        """Reference return value: sigmoid(fc(...)) [N,2] (the loss applies CE on it, Q4)."""
        return torch.sigmoid(self.forward_logits(x))

    def _init_weights(self):
        normal_init(self.conv1, 0, 0.01)
        normal_init(self.conv2, 0, 0.01)


# --------------------------------------------------------------------------------------
# NonLocalBlock and the heads built on it
# --------------------------------------------------------------------------------------
class NonLocalBlock(_HeadBase):
    """Attention over T tokens with the softmax taken over the QUERY axis (Q11).
    `forward_tokens` works on [T,C] token matrices (NHWC rows); `forward` keeps the reference's
    [b,C,h,w] signature."""

    BLOCK_TOKENS = 4096   # more tokens than this: functional.nonlocal_attention_blocked (the T x T scores are never materialised)
    BLOCK_K = 4096        # keys per block (measured at T = 32768: 512 -> 7.7 ms, 2048 -> 5.0 ms, 4096 -> 4.6 ms forward: launches amortise)

    def __init__(self, channel):
        super().__init__()
        self.inter_channel = channel // 2
        mk = lambda i, o: _channels_last_(nn.Conv2d(i, o, kernel_size=1, stride=1, padding=0, bias=False))
        self.conv_phi = mk(channel, self.inter_channel)
        self.conv_theta = mk(channel, self.inter_channel)
        self.conv_g = mk(channel, self.inter_channel)
        self.conv_mask = mk(self.inter_channel, channel)

    def _init_weights(self):
        for m in (self.conv_phi, self.conv_theta, self.conv_g, self.conv_mask):
            normal_init(m, 0, 0.01)

    def forward_tokens(self, x, grl=1.0):
        """x [T,C] (activation dtype) -> [T,C]; `grl` scales the gradient into x (first-layer fold)."""
        T, C = x.shape
        xin = x.view(T, 1, 1, C)
        lay = lambda conv, g: F_.dense_layer(xin, conv.weight, grl=g).view(T, -1)
        phi, theta, g = lay(self.conv_phi, grl), lay(self.conv_theta, grl), lay(self.conv_g, grl)
        if T > self.BLOCK_TOKENS:
            # key-block by key-block, two-pass query-axis softmax, recompute in backward: no T x T matrix
            y = F_.nonlocal_attention_blocked(theta, phi, g, self.BLOCK_K)
            mask = F_.dense_layer(y.view(T, 1, 1, -1), self.conv_mask.weight).view(T, C)
            skip = x if grl == 1.0 else F_.gradient_scalar(x, grl)
            return mask + skip
        # S[q,k] = theta[q,:].phi[k,:]  (theta as activations, phi as the "weight")
        s = F_.dense_layer(theta.view(T, 1, 1, -1), phi, out_dtype=torch.float32).view(T, T)
        p = F_.softmax_dim0(s)
        # Y[q,c] = sum_k P[q,k] g[k,c]  (g^T as the "weight" [C/2, T])
        y = F_.dense_layer(F_.cast(p, x.dtype).view(T, 1, 1, T), g.t().contiguous()).view(T, -1)
        mask = F_.dense_layer(y.view(T, 1, 1, -1), self.conv_mask.weight).view(T, C)
        # residual: identity path also carries the (possibly reversed) gradient
        skip = x if grl == 1.0 else F_.gradient_scalar(x, grl)
        return mask + skip

    def forward_single_tokens(self, x, grl=1.0):
        """Every row of x [T,C] as its own one-token sequence (the reference calls the block with [1,C] inputs from the
        per-RoI loops of group_local_da_loss, DAFaster_rcnn.py:311-313): the softmax over one score is 1, so
        y = x + W_mask W_g x and the phi / theta projections drop out."""
        T, C = x.shape
        g = F_.dense_layer(x.view(T, 1, 1, C), self.conv_g.weight, grl=grl)
        mask = F_.dense_layer(g, self.conv_mask.weight).view(T, C)
        skip = x if grl == 1.0 else F_.gradient_scalar(x, grl)
        return mask + skip

    def forward(self, x):
        b, c, h, w = x.shape
        a = F_.to_nhwc(x, F_.act_dtype())
        outs = [self.forward_tokens(a[i].reshape(h * w, c)) for i in range(b)]
        y = torch.stack(outs, 0).view(b, h, w, c)
        return F_.nhwc_to_nchw_view(y)


class NonLocalAlignmentHead(_HeadBase):
    """Deep backbone pixel head: GRL -> 1x1 conv + BN + ReLU + drop -> NonLocalBlock (C-channel out)."""

    def __init__(self, in_channels, context=False, grl=True):
        super().__init__()
        self.grl_flag = grl
        self.grl = GradientScalarLayer(-1.0)
        self.drop = nn.Dropout(p=0.5)
        self.conv1 = _channels_last_(nn.Conv2d(in_channels, in_channels, 1, bias=False))
        self.bn1 = nn.BatchNorm2d(in_channels)
        self.nlb1 = NonLocalBlock(int(in_channels))
        self.context = context
        self._init_weights()

    def _init_weights(self):
        self.nlb1._init_weights()
        normal_init(self.conv1, 0, 0.01)

    def forward(self, x):
        b, c, h, w = x.shape
        a = F_.to_nhwc(x, F_.act_dtype())
        g = self.grl.weight if self.grl_flag else 1.0
        t = self._layer(a, self.conv1, self.bn1, grl=g)
        outs = [self.nlb1.forward_tokens(t[i].reshape(h * w, c)) for i in range(b)]
        y = torch.stack(outs, 0).view(b, h, w, c)
        return F_.nhwc_to_nchw_view(y).float()


USE_CHAIN = True      # functional.instance_head_chain for the instance heads on the bf16 engine (False: layer by layer)
USE_FUSED_TAIL = True # functional.conv_pixel_loss for the pixel-level heads (False: separate pixel_head / loss kernels)


USE_CHAIN_FEED = True # let the chain kernel also run the FC that produces the RoI features (hotpath.SharedFCs.split)


def _chain_ok(x):
    return USE_CHAIN and x.is_cuda and x.dim() == 2 and x.shape[0] > 0 and F_.get_engine() == "umma_bf16" and x.shape[1] % 64 == 0 \
        and torch.is_grad_enabled()


def chain_feed_ok(xin):
    """The instance heads can take `pre=(xin, w0, b0, b_in)` (the layer producing their input runs inside the chain kernel)."""
    return USE_CHAIN_FEED and _chain_ok(xin) and xin.dtype == torch.bfloat16


class InstanceAlignmentHead(_HeadBase):
    """I1: GRL -> NonLocalBlock over the k RoIs -> FC 1024-512-512-2 -> sigmoid."""

    def __init__(self, context=False):
        super().__init__()
        self.grl = GradientScalarLayer(weight=-1.0)
        self.nlb = NonLocalBlock(1024)
        self.nlb._init_weights()
        self.fc1 = nn.Linear(1024, 512)
        self.bn1 = nn.BatchNorm1d(512)  # constructed but unused in the reference (Q9)
        self.fc2 = nn.Linear(512, 512)
        self.bn2 = nn.BatchNorm1d(512)  # unused
        self.fc3 = nn.Linear(512, 2)
        self.context = context

    def unused_parameters(self):
        return list(self.bn1.parameters()) + list(self.bn2.parameters())

    def forward_logits(self, x, single_token=False):
        """single_token: treat every row as a separate [1,1024] call of the reference head (no attention across rows)."""
        k = x.shape[0]
        a = F_.cast(x.contiguous(), F_.act_dtype())
        t = (self.nlb.forward_single_tokens if single_token else self.nlb.forward_tokens)(a, grl=self.grl.weight)
        t = self._layer(t.view(k, 1, 1, -1), self.fc1, relu=True, drop=True)
        t = self._layer(t, self.fc2, relu=True, drop=True)
        z = self._layer(t, self.fc3, relu=False, drop=False, engine=_tiny_engine(), out_dtype=torch.float32)
        return z.view(k, 2).float()

    def forward(self, x):
        return torch.sigmoid(self.forward_logits(x))

    def forward_loss(self, x, labels, pre=None):
        """(mean CE(sigmoid(fc3(...)), labels), pred = sigmoid(fc3(...))): the head AND the instance loss built on it
        (DAFaster_rcnn_Orig.py:177-188).  On the bf16 tensor-core engine this is ONE kernel forward and ONE backward
        (functional.instance_head_chain -> da_instance_fc_forward/backward); the other engines run layer by layer.
        pre=(xin, w0, b0, b_in) with x=None: the FC producing the features joins the kernel (see chain_feed_ok)."""
        if pre is not None or _chain_ok(x):
            p = self._p()
            seeds = (_draw_seed(self.training, p), _draw_seed(self.training, p))
            nlb = self.nlb
            return F_.instance_head_chain(None if pre is not None else F_.cast(x.contiguous(), torch.bfloat16), labels,
                                          (nlb.conv_theta.weight, nlb.conv_phi.weight, nlb.conv_g.weight), nlb.conv_mask.weight,
                                          ((self.fc1.weight, self.fc1.bias), (self.fc2.weight, self.fc2.bias), (self.fc3.weight, self.fc3.bias)),
                                          p, seeds, self.grl.weight, pre=pre)
        return F_.ce2(self.forward_logits(x), labels, True)

    def _init_weights(self):
        normal_init(self.fc1, 0, 0.01)
        normal_init(self.fc2, 0, 0.01)
        normal_init(self.fc3, 0, 0.05)


class InstanceAlignmentHead_DAF(_HeadBase):
    """I2: GRL -> FC 1024-1024-1024-2 -> sigmoid."""

    def __init__(self, context=False):
        super().__init__()
        self.grl = GradientScalarLayer(weight=-1.0)
        self.fc1 = nn.Linear(1024, 1024)
        self.fc2 = nn.Linear(1024, 1024)
        self.fc3 = nn.Linear(1024, 2)
        self.context = context

    def forward_logits(self, x):
        k = x.shape[0]
        a = F_.cast(x.contiguous(), F_.act_dtype())
        t = self._layer(a.view(k, 1, 1, -1), self.fc1, relu=True, drop=True, grl=self.grl.weight)
        t = self._layer(t, self.fc2, relu=True, drop=True)
        z = self._layer(t, self.fc3, relu=False, drop=False, engine=_tiny_engine(), out_dtype=torch.float32)
        return z.view(k, 2).float()

    def forward(self, x):
        return torch.sigmoid(self.forward_logits(x))

    def forward_loss(self, x, labels, pre=None):
        """See InstanceAlignmentHead.forward_loss (same kernel without the NonLocalBlock part)."""
        if pre is not None or _chain_ok(x):
            p = self._p()
            seeds = (_draw_seed(self.training, p), _draw_seed(self.training, p))
            return F_.instance_head_chain(None if pre is not None else F_.cast(x.contiguous(), torch.bfloat16), labels, None, None,
                                          ((self.fc1.weight, self.fc1.bias), (self.fc2.weight, self.fc2.bias), (self.fc3.weight, self.fc3.bias)),
                                          p, seeds, self.grl.weight, pre=pre)
        return F_.ce2(self.forward_logits(x), labels, True)

    def _init_weights(self):
        normal_init(self.fc1, 0, 0.01)
        normal_init(self.fc2, 0, 0.01)
        normal_init(self.fc3, 0, 0.01)
