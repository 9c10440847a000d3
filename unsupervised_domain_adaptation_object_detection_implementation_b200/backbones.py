"""DA backbones: a ResNet trunk (plain PyTorch/cuDNN — adjacent to the hot path) carrying the
image-level DA heads, with the reference's `forward_train(x, gt_domain)` return arities (§8b):

  ResNet_DAF       mmdet/models/backbones/resnet_da_daf_org.py:796-824   -> (outs, loss, img_feat)
  ResNet_DA        mmdet/models/backbones/resnet_da.py:821-850           -> (outs, loss_vec)
  ResNet_DA_CBAM   mmdet/models/backbones/resnet_da_cbam.py:934-993      -> (outs, loss_vec, patch_loss)
  ResNet_DA_Deep   mmdet/models/backbones/resnet_da_deep.py:1120-1175    -> (outs, loss_vec, patch_loss)

Trunk parameter names follow mmdet's ResNet (conv1, bn1, layer{1..4}.{i}.conv{1..3}/bn{1..3}/downsample.{0,1}),
DA-head attribute names follow the reference (Appendix C), so reference checkpoints load by key.  The loss
tails run in the fused loss kernels and stay on the device (the reference copies them into a CPU tensor, Q7)."""
import torch
import torch.nn as nn
from torch.nn.modules.batchnorm import _BatchNorm

from . import da_heads, da_losses
from .registry import BACKBONES


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        # style='pytorch': the stride sits on the 3x3 conv
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=stride, padding=dilation, dilation=dilation, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        identity = x if self.downsample is None else self.downsample(x)
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.relu(self.bn2(self.conv2(out)))
        out = self.bn3(self.conv3(out))
        return self.relu(out + identity)


class _ResNetTrunk(nn.Module):
    arch_settings = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3), 152: (3, 8, 36, 3)}

    def __init__(self, depth=50, in_channels=3, num_stages=4, strides=(1, 2, 2, 2), dilations=(1, 1, 1, 1),
                 out_indices=(0, 1, 2, 3), style="pytorch", frozen_stages=-1, norm_cfg=None, norm_eval=True,
                 init_cfg=None, pretrained=None, **kwargs):
        super().__init__()
        if depth not in self.arch_settings:
            raise KeyError(f"invalid depth {depth} for resnet")
        if style != "pytorch":
            raise NotImplementedError("the DA configs use style='pytorch'")
        self.depth, self.out_indices, self.frozen_stages, self.norm_eval = depth, tuple(out_indices), frozen_stages, norm_eval
        self.init_cfg = init_cfg
        self.deep_stem = False
        self.conv1 = nn.Conv2d(in_channels, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, stride=2, padding=1)
        self.res_layers = []
        inplanes = 64
        for i, nblocks in enumerate(self.arch_settings[depth][:num_stages]):
            planes = 64 * 2 ** i
            stride, dilation = strides[i], dilations[i]
            blocks = []
            down = None
            if stride != 1 or inplanes != planes * 4:
                down = nn.Sequential(nn.Conv2d(inplanes, planes * 4, 1, stride=stride, bias=False), nn.BatchNorm2d(planes * 4))
            blocks.append(Bottleneck(inplanes, planes, stride, dilation, down))
            inplanes = planes * 4
            for _ in range(1, nblocks):
                blocks.append(Bottleneck(inplanes, planes, 1, dilation))
            name = f"layer{i + 1}"
            self.add_module(name, nn.Sequential(*blocks))
            self.res_layers.append(name)
        self.feat_dim = inplanes
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        self._freeze_stages()

    @property
    def norm1(self):
        return self.bn1

    def _freeze_stages(self):
        if self.frozen_stages >= 0:
            self.bn1.eval()
            for m in (self.conv1, self.bn1):
                for p in m.parameters():
                    p.requires_grad = False
        for i in range(1, self.frozen_stages + 1):
            m = getattr(self, f"layer{i}")
            m.eval()
            for p in m.parameters():
                p.requires_grad = False

    def _stem(self, x):
        return self.maxpool(self.relu(self.bn1(self.conv1(x))))

    def _stages(self, x):
        x = self._stem(x)
        for i, name in enumerate(self.res_layers):
            x = getattr(self, name)(x)
            yield i, x

    def forward(self, x):
        return tuple(f for i, f in self._stages(x) if i in self.out_indices)

    def train(self, mode=True):
        """Keep normalisation layers frozen while training (resnet_da_cbam.py:995-1004)."""
        super().train(mode)
        self._freeze_stages()
        if mode and self.norm_eval:
            for m in self.modules():
                if isinstance(m, _BatchNorm):
                    m.eval()
        return self

    def unused_parameters(self):
        out = []
        for m in self.children():
            if hasattr(m, "unused_parameters"):
                out += m.unused_parameters()
        return out


@BACKBONES.register_module()
class ResNet_DAF(_ResNetTrunk):
    """DAF-Org: ImgAlignmentHead on C5 + whole-batch pixel loss L1 (Q5)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.da_head_top = da_heads.ImgAlignmentHead(2048)
        self.da_head_top._init_weights()
        self.da_outs_idx = (3,)

    def forward_train(self, x, gt_domain):
        outs, patch_feat, loss = [], None, None
        for i, f in self._stages(x):
            if i in self.da_outs_idx:
                if i in self.out_indices:
                    outs.append(f)
                if i == 3:
                    loss, patch_feat = self.da_head_top.forward_loss(f, gt_domain)      # H1 + L1: GEMM + fused tail kernel
        return tuple(outs), loss, patch_feat


@BACKBONES.register_module()
class ResNet_DA(_ResNetTrunk):
    """MAF: SRM heads on C3/C4/C5, CE on their sigmoid outputs (Q4)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.criterion = nn.CrossEntropyLoss()
        self.da_head_top = da_heads.SRM(2048)
        self.da_head_mid = da_heads.SRM(1024)
        self.da_head_bottom = da_heads.SRM(512)
        for h in (self.da_head_top, self.da_head_mid, self.da_head_bottom):
            h._init_weights()
        self.da_outs_idx = (1, 2, 3)

    def forward_train(self, x, gt_domain):
        outs, losses = [], []
        heads = {1: self.da_head_bottom, 2: self.da_head_mid, 3: self.da_head_top}
        for i, f in self._stages(x):
            if i in self.da_outs_idx:
                if i in self.out_indices:
                    outs.append(f)
                loss, _ = da_losses.image_ce_loss(heads[i].forward_logits(f), gt_domain, True)
                losses.append(loss)
        return tuple(outs), torch.stack(losses)


class _GlobalLocal(_ResNetTrunk):
    def _global_local(self, x, gt_domain, local_heads):
        outs, glob, patch = [], [], 0
        for i, f in self._stages(x):
            if i in self.da_outs_idx:
                if i in self.out_indices:
                    outs.append(f)
                if i in local_heads:
                    head = local_heads[i]
                    patch = patch + (head.forward_loss(f, gt_domain)[0] if hasattr(head, "forward_loss") else
                                     da_losses.patch_loss(head(f), gt_domain))
                if i == 2:
                    glob.append(da_losses.image_ce_loss(self.da_head_mid(f), gt_domain, False)[0])
                elif i == 3:
                    glob.append(da_losses.image_ce_loss(self.da_head_top(f), gt_domain, False)[0])
        return tuple(outs), torch.stack(glob), patch


@BACKBONES.register_module()
class ResNet_DA_CBAM(_GlobalLocal):
    """DAF (CBAM flavour): Global heads on C4/C5 (CE on raw logits) + Local head on C3 (per-image loss L2)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.patch_bottom_align, self.patch_mid_align = True, False
        self.criterion = nn.CrossEntropyLoss()
        self.da_head_top = da_heads.GlobalAlignmentHead(in_channel=2048, context=False)
        self.da_head_mid = da_heads.GlobalAlignmentHead(in_channel=1024, context=False, grl=True)
        self.local_da_head_bottom = da_heads.LocalAlignmentHead(in_channels=512, context=False)
        self.da_head_top._init_weights()
        self.da_head_mid._init_weights()
        self.da_outs_idx = (1, 2, 3)

    def forward_train(self, x, gt_domain):
        return self._global_local(x, gt_domain, {1: self.local_da_head_bottom})


@BACKBONES.register_module()
class ResNet_DA_Deep(_GlobalLocal):
    """DeepAlign: Global heads without the dead branch + NonLocalAlignmentHeads on C3/C4."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.patch_bottom_align, self.patch_mid_align = True, True
        self.criterion = nn.CrossEntropyLoss()
        self.da_head_top = da_heads.GlobalAlignmentHeadDeep(in_channel=2048, context=False)
        self.da_head_mid = da_heads.GlobalAlignmentHeadDeep(in_channel=1024, context=False, grl=True)
        self.local_da_head_mid = da_heads.NonLocalAlignmentHead(in_channels=1024, context=False, grl=True)
        self.local_da_head_bottom = da_heads.NonLocalAlignmentHead(in_channels=512, context=False)
        self.da_head_top._init_weights()
        self.da_head_mid._init_weights()
        self.da_outs_idx = (1, 2, 3)

    def forward_train(self, x, gt_domain):
        return self._global_local(x, gt_domain, {1: self.local_da_head_bottom, 2: self.local_da_head_mid})
