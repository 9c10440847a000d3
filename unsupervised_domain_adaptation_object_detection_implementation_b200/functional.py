"""torch.autograd bindings of the C-ABI kernels (libda_b200.so).

PyTorch is plumbing here: it owns device memory, streams and the autograd tape; every
forward/backward below is ONE or a few calls into the library on the current CUDA stream.
Tensors that are not on a CUDA device raise: there is no CPU fallback (the CPU restatement
lives in oracle/ and is test infrastructure only).

Activation layout inside the DA heads is NHWC ([N,H,W,C] contiguous); `to_nhwc` accepts the
reference's NCHW tensors (zero-copy when they are channels_last).
"""
import ctypes
import math

import os
import torch
from torch.autograd import Function

from . import _lib
from ._lib import lib, check

_WS = {}


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("libda_b200 kernels need CUDA tensors (sm_100a); there is no CPU fallback")


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _code(dtype):
    if dtype == torch.float32:
        return _lib.DA_F32
    if dtype == torch.bfloat16:
        return _lib.DA_BF16
    raise RuntimeError(f"unsupported dtype {dtype} (float32 / bfloat16 only)")


def workspace(nbytes, device, tag="default"):
    """Caller-owned scratch (the library never allocates).  Grown on demand, reused."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), tag)
    buf = _WS.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1024), dtype=torch.uint8, device=device)
        _WS[key] = buf
    return buf


# --------------------------------------------------------------------------------------
# engine selection
# --------------------------------------------------------------------------------------
_ENGINE = ["umma_bf16"]


def set_sm_limit(n):
    """SM budget of the persistent kernels for the launches that follow (0 = all SMs); see da_set_sm_limit."""
    check(lib.da_set_sm_limit(int(n)), "set_sm_limit")


def set_option(name, value):
    """Debug / test switch of the library (da_set_option): e.g. set_option("roi_no_tc", 1) forces the CUDA-core RoIAlign
    kernels for bf16 tensors.  The DA_* environment variables are read once, when the library is loaded."""
    check(lib.da_set_option(name.encode(), int(value)), "set_option")
    _OPTIONS[name] = int(value)


_OPTIONS = {"roi_no_tc": 1 if os.environ.get("DA_ROI_NO_TC") else 0}


def set_engine(name):
    """'umma_bf16' (tcgen05, default), 'umma_bf16x6' (fp32-class tcgen05: exact 3-way bf16 split, six product
    terms, fp32 in/out, <= 1e-5), 'umma_bf16x3' (2-way split, ~2^-16) or 'simt_f32' (CUDA-core fp32 parity engine)."""
    if name not in _lib.ENGINES:
        raise ValueError(f"unknown engine {name!r}; choose from {sorted(_lib.ENGINES)}")
    _ENGINE[0] = name


def get_engine():
    return _ENGINE[0]


def act_dtype(engine=None):
    return torch.bfloat16 if (engine or _ENGINE[0]) == "umma_bf16" else torch.float32


# --------------------------------------------------------------------------------------
# layout
# --------------------------------------------------------------------------------------
class _NCHWtoNHWC(Function):
    @staticmethod
    def forward(ctx, x, dtype):
        _require_cuda(x)
        N, C, H, W = x.shape
        x = x.contiguous()
        out = torch.empty((N, H, W, C), dtype=dtype, device=x.device)
        check(lib.da_nchw_to_nhwc(_ptr(x), _code(x.dtype), _ptr(out), _code(dtype), N, C, H, W, _stream()), "nchw_to_nhwc")
        ctx.src_dtype = x.dtype
        return out

    @staticmethod
    def backward(ctx, g):
        N, H, W, C = g.shape
        g = g.contiguous()
        out = torch.empty((N, C, H, W), dtype=ctx.src_dtype, device=g.device)
        check(lib.da_nhwc_to_nchw(_ptr(g), _code(g.dtype), _ptr(out), _code(out.dtype), N, C, H, W, _stream()), "nhwc_to_nchw")
        return out, None


class _Cast(Function):
    @staticmethod
    def forward(ctx, x, dtype):
        ctx.src_dtype = x.dtype
        x = x.contiguous()
        out = torch.empty(x.shape, dtype=dtype, device=x.device)
        check(lib.da_cast(_ptr(x), _code(x.dtype), _ptr(out), _code(dtype), x.numel(), _stream()), "cast")
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        out = torch.empty(g.shape, dtype=ctx.src_dtype, device=g.device)
        check(lib.da_cast(_ptr(g), _code(g.dtype), _ptr(out), _code(out.dtype), g.numel(), _stream()), "cast")
        return out, None


def cast(x, dtype):
    return x if x.dtype == dtype else _Cast.apply(x, dtype)


def to_nhwc(x, dtype=None):
    """NCHW (logical) tensor -> contiguous [N,H,W,C] in `dtype`.  Zero-copy for channels_last."""
    _require_cuda(x)
    dtype = dtype or x.dtype
    if x.dim() != 4:
        raise RuntimeError("to_nhwc expects a 4-D NCHW tensor")
    v = x.permute(0, 2, 3, 1)
    if v.is_contiguous():  # channels_last storage: NHWC is a view
        return cast(v, dtype)
    return _NCHWtoNHWC.apply(x, dtype)


def nhwc_to_nchw_view(y):
    """[N,H,W,C] -> logical NCHW view (channels_last strides, no copy)."""
    return y.permute(0, 3, 1, 2)


class _GRL(Function):
    """mmdet/models/roi_heads/instance_da.py:14-23 (_GradientScalarLayer)."""

    @staticmethod
    def forward(ctx, x, weight):
        ctx.weight = float(weight)
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        _require_cuda(g)
        g = g.contiguous()
        out = torch.empty_like(g)
        check(lib.da_grl_backward(_ptr(g), _ptr(out), _code(g.dtype), g.numel(), ctx.weight, _stream()), "grl_backward")
        return out, None


def gradient_scalar(x, weight):
    return _GRL.apply(x, weight)


# --------------------------------------------------------------------------------------
# RoIAlign
# --------------------------------------------------------------------------------------
_ROI_PREP = {}      # device index -> what the shared "roi" workspace currently holds the preparation of


def _roi_prep_token(ws, rois_c, R, N, H, W, scale, sr, aligned):
    return (ws.data_ptr(), rois_c.data_ptr(), rois_c._version, R, N, H, W, float(scale), int(sr), int(aligned), _stream().value)


class RoIAlignFunction(Function):
    """mmcv.ops.roi_align.RoIAlignFunction replacement (avg pooling, 7x7)."""

    @staticmethod
    def forward(ctx, feat, rois, output_size, spatial_scale, sampling_ratio, aligned, out_dtype=None,
                out_layout="rchw", validate=False):
        _require_cuda(feat, rois)
        if feat.dim() != 4 or rois.dim() != 2 or rois.shape[1] != 5:
            raise RuntimeError("roi_align: feat must be [N,C,H,W] and rois [R,5]")
        ph, pw = (output_size, output_size) if isinstance(output_size, int) else tuple(output_size)
        N, C, H, W = feat.shape
        R = rois.shape[0]
        out_dtype = out_dtype or feat.dtype
        f = to_nhwc(feat.detach())
        rois_c = rois.detach().to(torch.float32).contiguous()
        layout = _lib.ROI_OUT_RCHW if out_layout == "rchw" else _lib.ROI_OUT_RHWC
        shape = (R, C, ph, pw) if out_layout == "rchw" else (R, ph, pw, C)
        out = torch.empty(shape, dtype=out_dtype, device=feat.device)
        grid = torch.empty((R, 2), dtype=torch.int32, device=feat.device)
        nbytes = lib.da_roi_align_workspace_bytes(R, H, W)
        ws = workspace(nbytes, feat.device, "roi")
        check(lib.da_roi_align_forward(_ptr(f), _code(f.dtype), N, C, H, W, _ptr(rois_c), R, ph, pw,
                                       float(spatial_scale), int(sampling_ratio), int(bool(aligned)),
                                       _ptr(out), _code(out_dtype), layout, _ptr(grid), _ptr(ws), ws.numel(),
                                       _stream()), "roi_align_forward")
        _ROI_PREP[feat.device.index] = _roi_prep_token(ws, rois_c, R, N, H, W, spatial_scale, sampling_ratio, bool(aligned))
        if validate and R > 0:
            flag = ws[:4].view(torch.int32)[0].item()
            if flag != 0:
                raise RuntimeError("roi_align: a RoI carries a batch index outside [0, N) (reference Q1)")
        ctx.save_for_backward(rois_c)
        ctx.cfg = (N, C, H, W, ph, pw, float(spatial_scale), int(sampling_ratio), int(bool(aligned)), layout, feat.dtype)
        ctx.sampling_grid = grid
        ctx.mark_non_differentiable(grid)
        return out, grid

    @staticmethod
    def backward(ctx, gout, _ggrid):
        (rois_c,) = ctx.saved_tensors
        N, C, H, W, ph, pw, scale, sr, aligned, layout, fdtype = ctx.cfg
        R = rois_c.shape[0]
        gout = gout.contiguous()
        # bf16 features + bf16 [R,C,7,7] gradients: the tensor-core kernel rounds its fp32 accumulators once and
        # writes bf16 directly (no fp32 staging tensor, no cast pass)
        # ([R,7,7,C] gradients: the kernel's TMA fetches the operand as it lies in memory; needs C % 64 == 0)
        direct = fdtype == torch.bfloat16 and gout.dtype == torch.bfloat16 and C % (8 if layout == 0 else 64) == 0 and R > 0 \
            and not _OPTIONS.get("roi_no_tc")
        gin = torch.empty((N, H, W, C), dtype=torch.bfloat16 if direct else torch.float32, device=gout.device)
        ws = workspace(lib.da_roi_align_workspace_bytes(R, H, W), gout.device, "roi")
        # the forward's preparation (tap tables, footprints) is still in the workspace unless another roi_align call used it since
        token = _roi_prep_token(ws, rois_c, R, N, H, W, scale, sr, aligned)
        fn = lib.da_roi_align_backward_prepared if (R > 0 and _ROI_PREP.get(gout.device.index) == token) else lib.da_roi_align_backward
        _ROI_PREP[gout.device.index] = token
        check(fn(_ptr(gout), _code(gout.dtype), layout, _ptr(rois_c), R, ph, pw, scale, sr,
                 aligned, _ptr(gin), _code(gin.dtype), N, C, H, W, _ptr(ws), ws.numel(), _stream()),
              "roi_align_backward")
        g = nhwc_to_nchw_view(gin)
        if fdtype != torch.float32:
            g = g.to(fdtype)
        return g, None, None, None, None, None, None, None, None


def roi_align(feat, rois, output_size=7, spatial_scale=1.0, sampling_ratio=0, aligned=True, out_dtype=None,
              out_layout="rchw", validate=False, return_grid=False):
    out, grid = RoIAlignFunction.apply(feat, rois, output_size, spatial_scale, sampling_ratio, aligned, out_dtype,
                                       out_layout, validate)
    return (out, grid) if return_grid else out


def map_roi_levels(rois, num_levels, finest_scale=56.0):
    """single_level_roi_extractor.py:36-55."""
    _require_cuda(rois)
    rois_c = rois.detach().to(torch.float32).contiguous()
    out = torch.empty((rois_c.shape[0],), dtype=torch.int32, device=rois.device)
    check(lib.da_map_roi_levels(_ptr(rois_c), rois_c.shape[0], int(num_levels), float(finest_scale), _ptr(out), _stream()),
          "map_roi_levels")
    return out.long()


# --------------------------------------------------------------------------------------
# losses
# --------------------------------------------------------------------------------------
def _as_i32(t, device):
    if not torch.is_tensor(t):
        t = torch.as_tensor(t, device=device)
    return t.to(device=device, dtype=torch.int32).contiguous()


class PixelDomainLossFunction(Function):
    """L1/L2: 0.5*mean(sigmoid(p)^2) (source) / 0.5*mean(sigmoid(1-p)^2) (target)."""

    @staticmethod
    def forward(ctx, logits, domain, whole_batch):
        _require_cuda(logits)
        x = logits.contiguous().float()
        N = x.shape[0]
        Lp = x.numel() // N
        dom = _as_i32(domain, x.device)
        loss = torch.empty((), dtype=torch.float32, device=x.device)
        ws = workspace(lib.da_pixel_loss_workspace_bytes(N, Lp), x.device, "loss")
        check(lib.da_pixel_domain_loss_forward(_ptr(x), N, Lp, _ptr(dom), int(whole_batch), _ptr(loss), _ptr(ws),
                                               ws.numel(), _stream()), "pixel_domain_loss_forward")
        ctx.save_for_backward(x, dom)
        ctx.whole = int(whole_batch)
        ctx.shape = logits.shape
        return loss

    @staticmethod
    def backward(ctx, g):
        x, dom = ctx.saved_tensors
        N = x.shape[0]
        Lp = x.numel() // N
        g = g.contiguous().float()
        dx = torch.empty_like(x)
        check(lib.da_pixel_domain_loss_backward(_ptr(x), N, Lp, _ptr(dom), ctx.whole, _ptr(g), 1.0, _ptr(dx), _stream()),
              "pixel_domain_loss_backward")
        return dx.view(ctx.shape), None, None


def pixel_domain_loss(logits, domain, whole_batch):
    return PixelDomainLossFunction.apply(logits, domain, whole_batch)


class CE2Function(Function):
    """nn.CrossEntropyLoss over 2 classes on raw logits or on sigmoid outputs (Q4).
    Returns (loss, pred) where pred = sigmoid(z) (what the reference heads return)."""

    @staticmethod
    def forward(ctx, z, labels, on_sigmoid):
        _require_cuda(z)
        zc = z.contiguous().float()
        R = zc.shape[0]
        lab = _as_i32(labels, zc.device)
        loss = torch.empty((), dtype=torch.float32, device=zc.device)
        pred = torch.empty_like(zc)
        check(lib.da_ce2_forward(_ptr(zc), _ptr(lab), R, int(on_sigmoid), _ptr(pred), _ptr(loss), _stream()), "ce2_forward")
        ctx.save_for_backward(zc, lab)
        ctx.on_sigmoid = int(on_sigmoid)
        return loss, pred

    @staticmethod
    def backward(ctx, gloss, gpred):
        zc, lab = ctx.saved_tensors
        R = zc.shape[0]
        dz = torch.empty_like(zc)
        gl = gloss.contiguous().float()
        gp = gpred.contiguous().float() if (gpred is not None and ctx.on_sigmoid) else None
        check(lib.da_ce2_backward(_ptr(zc), _ptr(lab), R, ctx.on_sigmoid, _ptr(gl), 1.0, _ptr(gp), _ptr(dz), _stream()),
              "ce2_backward")
        return dz, None, None


def ce2(z, labels, on_sigmoid):
    return CE2Function.apply(z, labels, on_sigmoid)


class Focal2Function(Function):
    """mmcv sigmoid_focal_loss on [k,2] with mean reduction (losses/focal_loss.py:60-103)."""

    @staticmethod
    def forward(ctx, u, labels, gamma, alpha):
        _require_cuda(u)
        uc = u.contiguous().float()
        k = uc.shape[0]
        lab = _as_i32(labels, uc.device)
        loss = torch.empty((), dtype=torch.float32, device=uc.device)
        check(lib.da_focal2_forward(_ptr(uc), _ptr(lab), k, float(gamma), float(alpha), _ptr(loss), _stream()), "focal2_forward")
        ctx.save_for_backward(uc, lab)
        ctx.cfg = (float(gamma), float(alpha))
        return loss

    @staticmethod
    def backward(ctx, g):
        uc, lab = ctx.saved_tensors
        du = torch.empty_like(uc)
        g = g.contiguous().float()
        check(lib.da_focal2_backward(_ptr(uc), _ptr(lab), uc.shape[0], ctx.cfg[0], ctx.cfg[1], _ptr(g), 1.0, _ptr(du),
                                     _stream()), "focal2_backward")
        return du, None, None, None


def sigmoid_focal_loss2(u, labels, gamma=2.0, alpha=0.25):
    return Focal2Function.apply(u, labels, gamma, alpha)


class ConsistencyFunction(Function):
    """DAFaster_rcnn_Orig.py:161-175 in closed form: sum_r |mean(sigmoid(img)) - sigmoid(pred[r,label_r])|."""

    @staticmethod
    def forward(ctx, img_logits, ins_pred, labels):
        _require_cuda(img_logits, ins_pred)
        a = img_logits.contiguous().float()
        p = ins_pred.contiguous().float()
        lab = _as_i32(labels, a.device)
        loss = torch.empty((), dtype=torch.float32, device=a.device)
        mean = torch.empty((), dtype=torch.float32, device=a.device)
        check(lib.da_consistency_forward(_ptr(a), a.numel(), _ptr(p), _ptr(lab), p.shape[0], _ptr(mean), _ptr(loss),
                                         _stream()), "consistency_forward")
        ctx.save_for_backward(a, p, lab, mean)
        ctx.img_shape = img_logits.shape
        return loss

    @staticmethod
    def backward(ctx, g):
        a, p, lab, mean = ctx.saved_tensors
        g = g.contiguous().float()
        da_ = torch.empty_like(a)
        dp = torch.empty_like(p)
        check(lib.da_consistency_backward(_ptr(a), a.numel(), _ptr(p), _ptr(lab), p.shape[0], _ptr(mean), _ptr(g), 1.0,
                                          _ptr(da_), _ptr(dp), _stream()), "consistency_backward")
        return da_.view(ctx.img_shape), dp, None


def consistency_loss(img_logits, ins_pred, labels):
    return ConsistencyFunction.apply(img_logits, ins_pred, labels)


# --------------------------------------------------------------------------------------
# dense layers (conv / linear) with fused epilogue
# --------------------------------------------------------------------------------------
def _conv_desc(N, H, W, Cin, Cout, KH, KW, stride, pad, engine, x_dtype, y_dtype):
    return _lib.ConvDesc(N, H, W, Cin, Cout, KH, KW, stride, pad, _lib.ENGINES[engine], _code(x_dtype), _code(y_dtype))


def bf16_shadow(w, refresh=False):
    """bf16 copy of an fp32 weight, cached ON the parameter object together with the version it was
    made from.  FusedSGD refreshes the copy inside its update kernel (without bumping the version),
    so the steady-state train step never re-casts; any torch in-place update invalidates it."""
    hit = getattr(w, "_da_shadow", None)
    if hit is not None and not refresh and hit[0] == w._version and hit[1].shape == w.shape and hit[1].device == w.device:
        return hit[1]
    sh = torch.empty_like(w, dtype=torch.bfloat16)          # preserves strides (channels_last weights)
    src = w.detach()
    check(lib.da_cast(_ptr(src), _lib.DA_F32, _ptr(sh), _lib.DA_BF16, src.numel(), _stream()), "cast")
    w._da_shadow = (w._version, sh)
    return sh


def _dense_memory(t):
    """True when the tensor occupies one dense block (any permutation of dims)."""
    return t.is_contiguous() or (t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last))


def _w_ohwi(w):
    """Parameter [O,I,KH,KW] (or [O,I]) -> contiguous [O,KH,KW,I] view (zero-copy when the
    parameter is stored channels_last, which DenseConv does at construction)."""
    if w.dim() == 2:
        return w.contiguous().view(w.shape[0], 1, 1, w.shape[1])
    v = w.permute(0, 2, 3, 1)
    return v if v.is_contiguous() else v.contiguous()


# callable(dw) invoked right after a layer's weight-gradient kernel has been enqueued (None = disabled)
WGRAD_HOOK = None
# weight gradient of small layers on a side stream, next to the data gradient (see DenseLayerFunction.backward)
CONCURRENT_SMALL_WGRAD = not os.environ.get("DA_NO_CONCURRENT_WGRAD")
SMALL_LAYER_FLOPS = 6e9
_SIDE = {}


def _side_stream(dev):
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    st = _SIDE.get(key)
    if st is None:
        st = _SIDE[key] = torch.cuda.Stream(device=dev)
    return st
# Weights whose gradient never becomes a .grad tensor: id(weight) -> ManagedWeight.  Autograd sees such a weight as a
# constant; the layer's backward hands the gradient to the optimizer that registered it:
#   * peer.PeerShardedSGD (N GPUs): the weight-gradient kernel writes `grad`, `after_wgrad` / `layer_done` start the
#     exchange and the sharded update on a side stream;
#   * optim.FusedSGD(fuse_wgrad=...) (one GPU): `fuse()` returns the da_sgd_fuse record of this call and the
#     weight-gradient kernel applies the update in its epilogue (da_conv_backward_weight_sgd); no gradient is stored.
MANAGED_WGRAD = {}


class ManagedWeight:
    grad = None        # flat fp32 buffer for the weight gradient (None with `fuse`)
    fuse = None        # callable -> _lib.SgdFuse

    def before_forward(self):  # in front of the layer's forward kernel (it reads the operand copy)
        pass

    def after_wgrad(self):     # the weight-gradient kernel is enqueued
        pass

    def layer_done(self):      # every backward kernel of the layer is enqueued (it no longer reads its operand copy)
        pass


class DenseLayerFunction(Function):
    """y = dropout(relu(conv(x, w) * scale + shift)), NHWC in / NHWC out.

    Backward: activation derivative + dgrad (times `grl`, the gradient-reversal weight when x
    is a head input) + wgrad + the two column sums that give d(scale), d(shift)."""

    @staticmethod
    def forward(ctx, x, w, scale, shift, stride, pad, relu, drop_p, seed, engine, grl, out_dtype, shadow=None, managed=None,
                preact_grad=False):
        _require_cuda(x, w)
        ctx.managed = managed
        ctx.preact_grad = preact_grad
        N, H, W_, Cin = x.shape
        x = x.contiguous()
        if shadow is not None:
            wv = _w_ohwi(shadow)
        else:
            wv = _w_ohwi(w.detach())
            if wv.dtype != x.dtype:
                wv = cast(wv, x.dtype)
        Cout, KH, KW, _ = wv.shape
        out_dtype = out_dtype or x.dtype
        OH = (H + 2 * pad - KH) // stride + 1
        OW = (W_ + 2 * pad - KW) // stride + 1
        y = torch.empty((N, OH, OW, Cout), dtype=out_dtype, device=x.device)
        desc = _conv_desc(N, H, W_, Cin, Cout, KH, KW, stride, pad, engine, x.dtype, out_dtype)
        ws = workspace(lib.da_conv_workspace_bytes(ctypes.byref(desc)), x.device, "conv")
        sc = None if scale is None else scale.detach().float().contiguous()
        sh = None if shift is None else shift.detach().float().contiguous()
        check(lib.da_conv_forward(ctypes.byref(desc), _ptr(x), _ptr(wv), _ptr(sc), _ptr(sh), int(relu), float(drop_p),
                                  int(seed), _ptr(y), _ptr(ws), ws.numel(), _stream()), "conv_forward")
        ctx.save_for_backward(x, wv, y, sc, sh)
        ctx.cfg = (stride, pad, int(relu), float(drop_p), int(seed), engine, grl, w.shape, w.dtype, out_dtype)
        ctx.has = (scale is not None, shift is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, wv, y, sc, sh = ctx.saved_tensors
        stride, pad, relu, drop_p, seed, engine, grl, wshape, wdtype, out_dtype = ctx.cfg
        N, H, W_, Cin = x.shape
        Cout, KH, KW, _ = wv.shape
        dev = x.device
        # dz in the operand dtype of this layer
        desc_a = _conv_desc(N, H, W_, Cin, Cout, KH, KW, stride, pad, engine, out_dtype, out_dtype)
        ws = workspace(lib.da_conv_workspace_bytes(ctypes.byref(desc_a)), dev, "conv")
        dy = cast(dy.contiguous(), out_dtype)
        need_scale, need_shift = ctx.needs_input_grad[2], ctx.needs_input_grad[3]
        trivial = (not relu) and drop_p == 0.0 and sc is None and not (need_scale or need_shift)
        if ctx.preact_grad:
            # the consumer (functional.instance_head_chain with a feeding layer) already applied this layer's ReLU derivative
            # and produced the bias gradient: dy IS dz
            if need_scale or need_shift or sc is not None or drop_p != 0.0:
                raise RuntimeError("dense_layer(preact_grad=True): pass a detached shift, no scale, no dropout")
            trivial = True
        dshift = dvdot = None
        if trivial:
            dz = dy
        else:
            dz = torch.empty_like(dy)
            if need_scale or need_shift:
                dshift = torch.empty((Cout,), dtype=torch.float32, device=dev)
            if need_scale:
                dvdot = torch.empty((Cout,), dtype=torch.float32, device=dev)
            check(lib.da_conv_act_backward(ctypes.byref(desc_a), _ptr(dy), _ptr(y), _ptr(sc), relu, drop_p, seed, _ptr(dz),
                                           _ptr(dshift), _ptr(dvdot), _ptr(ws), ws.numel(), _stream()), "conv_act_backward")
        dz = cast(dz, x.dtype)
        dx = dw = dscale = None
        # weight gradient first: a gradient all-reduce can then start while the data gradient of the same layer runs
        # (WGRAD_HOOK, installed by dist.OverlappedGradAllReduce)
        managed = ctx.managed
        if managed is not None and managed.fuse is None:
            desc_w = _conv_desc(N, H, W_, Cin, Cout, KH, KW, stride, pad, engine, x.dtype, x.dtype)
            check(lib.da_conv_backward_weight(ctypes.byref(desc_w), _ptr(x), _ptr(dz), _ptr(managed.grad), _ptr(ws), ws.numel(),
                                              _stream()), "conv_backward_weight")
            managed.after_wgrad()     # the gradient buffer is complete in stream order
        elif managed is None and ctx.needs_input_grad[1]:
            desc_w = _conv_desc(N, H, W_, Cin, Cout, KH, KW, stride, pad, engine, x.dtype, x.dtype)
            dwv = torch.empty((Cout, KH, KW, Cin), dtype=torch.float32, device=dev)
            # Small layers (the 1-2 GFLOP GEMMs of the instance head fill less than half of the SMs each and cost a fixed
            # ~10 us of latency): weight and data gradient only share dz, so the weight gradient goes to a side stream and
            # the two run next to each other; joined before this function returns.
            flops = 2.0 * dz.numel() * KH * KW * Cin
            join = None
            if (CONCURRENT_SMALL_WGRAD and ctx.needs_input_grad[0] and engine == "umma_bf16" and wdtype == torch.float32
                    and flops <= SMALL_LAYER_FLOPS):
                side = _side_stream(dev)
                fork = torch.cuda.Event()
                fork.record()
                side.wait_event(fork)
                ws_side = workspace(lib.da_conv_workspace_bytes(ctypes.byref(desc_w)), dev, "conv_side")
                with torch.cuda.stream(side):
                    check(lib.da_conv_backward_weight(ctypes.byref(desc_w), _ptr(x), _ptr(dz), _ptr(dwv), _ptr(ws_side), ws_side.numel(),
                                                      _stream()), "conv_backward_weight")
                    if WGRAD_HOOK is not None:
                        WGRAD_HOOK(dwv)
                    join = torch.cuda.Event()
                    join.record()
            else:
                check(lib.da_conv_backward_weight(ctypes.byref(desc_w), _ptr(x), _ptr(dz), _ptr(dwv), _ptr(ws), ws.numel(),
                                                  _stream()), "conv_backward_weight")
            dw = dwv.view(wshape) if len(wshape) == 2 else dwv.permute(0, 3, 1, 2)
            if wdtype != torch.float32:
                dw = dw.to(wdtype)
            elif WGRAD_HOOK is not None and join is None:
                WGRAD_HOOK(dwv)
        if ctx.needs_input_grad[0]:
            desc_d = _conv_desc(N, H, W_, Cin, Cout, KH, KW, stride, pad, engine, x.dtype, x.dtype)
            dx = torch.empty_like(x)
            check(lib.da_conv_backward_data(ctypes.byref(desc_d), _ptr(dz), _ptr(wv), float(grl), _ptr(dx), _ptr(ws),
                                            ws.numel(), _stream()), "conv_backward_data")
        if managed is None and ctx.needs_input_grad[1] and join is not None:
            torch.cuda.current_stream().wait_event(join)
        if managed is not None:
            if managed.fuse is not None:     # data gradient first: the fused kernel rewrites the operand copy it read
                desc_w = _conv_desc(N, H, W_, Cin, Cout, KH, KW, stride, pad, engine, x.dtype, x.dtype)
                rec = managed.fuse()
                check(lib.da_conv_backward_weight_sgd(ctypes.byref(desc_w), _ptr(x), _ptr(dz), ctypes.byref(rec), _ptr(ws),
                                                      ws.numel(), _stream()), "conv_backward_weight_sgd")
            managed.layer_done()
        if need_scale:
            # v = acc*scale + shift  =>  d(scale) = sum dv*acc = (dvdot - shift*dshift) / scale
            t = sh if sh is not None else torch.zeros_like(dshift)
            safe = torch.where(sc == 0, torch.ones_like(sc), sc)
            dscale = torch.where(sc == 0, torch.zeros_like(sc), (dvdot - t * dshift) / safe)
        return dx, dw, dscale, (dshift if need_shift else None), None, None, None, None, None, None, None, None, None, None, None


def _umma_ok(x, w):
    cin = x.shape[-1]
    cout = w.shape[0]
    taps = 1 if w.dim() == 2 else w.shape[2] * w.shape[3]
    return cin % 8 == 0 and cout % 8 == 0 and taps <= 16


def dense_layer(x, w, scale=None, shift=None, stride=1, pad=0, relu=False, drop_p=0.0, seed=0, engine=None,
                grl=1.0, out_dtype=None, preact_grad=False):
    """preact_grad=True: the gradient that will arrive for the output is already the PRE-activation gradient (its consumer
    applied relu' and owns the bias gradient): backward skips the activation pass; `shift` must not require grad."""
    engine = engine or get_engine()
    if engine != "simt_f32" and not (_umma_ok(x, w) and stride <= 2):
        engine = "simt_f32"  # shapes the tensor-core tiles cannot express (e.g. the 2-logit FC)
    shadow = None
    if w.dtype == torch.float32 and x.dtype == torch.bfloat16 and w.is_leaf and _dense_memory(w):
        shadow = bf16_shadow(w)      # cached on the parameter; refreshed by FusedSGD
    managed = MANAGED_WGRAD.get(id(w)) if (MANAGED_WGRAD and torch.is_grad_enabled()) else None
    if managed is not None:
        if shadow is None:
            raise RuntimeError("a peer-managed weight needs the bf16 tensor-core engine (its operand copy is what the peers refresh)")
        w = w.detach()
        managed.before_forward()
    return DenseLayerFunction.apply(x, w, scale, shift, stride, pad, relu, drop_p, seed, engine, grl, out_dtype, shadow, managed,
                                    preact_grad)


_DROP_COUNTER = {}


def dropout_counter(device):
    """Device-resident uint64 step counter that every dropout kernel adds to its seed at run time.
    `bump_dropout_counter` (one tiny device op, graph-capturable) advances it once per train step, so a
    CUDA-graph replay draws fresh masks although the per-layer seeds are baked into the graph."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    t = _DROP_COUNTER.get(idx)
    if t is None:
        t = torch.zeros(1, dtype=torch.int64, device=device)
        _DROP_COUNTER[idx] = t
        check(lib.da_set_dropout_counter(_ptr(t)), "set_dropout_counter")
    return t


def bump_dropout_counter(device):
    dropout_counter(device).add_(0x9E3779B97F4A7C15 >> 1)


def dropout_keep_mask(seed, shape, drop_p, device):
    """The exact keep-mask the fused epilogues use (exported for the oracle)."""
    n = 1
    for s in shape:
        n *= int(s)
    out = torch.empty((n,), dtype=torch.uint8, device=device)
    check(lib.da_dropout_mask(int(seed), n, float(drop_p), _ptr(out), _stream()), "dropout_mask")
    return out.view(*shape).bool()


class PixelHeadFunction(Function):
    """Terminal 1-channel conv (+bias)(+ReLU): logits[m] = act(x[m,:].w + b)."""

    @staticmethod
    def forward(ctx, x, w, bias, relu):
        _require_cuda(x, w)
        x = x.contiguous()
        K = x.shape[-1]
        M = x.numel() // K
        wv = w.detach().reshape(-1).float().contiguous()
        bv = None if bias is None else bias.detach().reshape(-1).float().contiguous()
        logits = torch.empty(x.shape[:-1], dtype=torch.float32, device=x.device)
        check(lib.da_pixel_head_forward(_ptr(x), _code(x.dtype), M, K, _ptr(wv), _ptr(bv), int(relu), _ptr(logits), _stream()),
              "pixel_head_forward")
        ctx.save_for_backward(x, wv, logits)
        ctx.cfg = (int(relu), w.shape, bias is not None)
        return logits

    @staticmethod
    def backward(ctx, g):
        x, wv, logits = ctx.saved_tensors
        relu, wshape, has_bias = ctx.cfg
        K = x.shape[-1]
        M = x.numel() // K
        g = g.contiguous().float()
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw = torch.empty((K,), dtype=torch.float32, device=x.device)
        db = torch.empty((1,), dtype=torch.float32, device=x.device) if has_bias else None
        ws = workspace(lib.da_pixel_head_workspace_bytes(M, K), x.device, "head")
        check(lib.da_pixel_head_backward(_ptr(x), _code(x.dtype), M, K, _ptr(wv), _ptr(g), _ptr(logits) if relu else None,
                                         _ptr(dx), _code(x.dtype), _ptr(dw), _ptr(db), _ptr(ws), ws.numel(), _stream()),
              "pixel_head_backward")
        return dx, dw.view(wshape), db, None


def pixel_head(x, w, bias=None, relu=False):
    return PixelHeadFunction.apply(x, w, bias, relu)


class GlobalAvgPoolFunction(Function):
    @staticmethod
    def forward(ctx, x):
        _require_cuda(x)
        x = x.contiguous()
        N, H, W_, C = x.shape
        y = torch.empty((N, C), dtype=torch.float32, device=x.device)
        ws = workspace(lib.da_global_avgpool_workspace_bytes(N, C), x.device, "pool")
        check(lib.da_global_avgpool_forward(_ptr(x), _code(x.dtype), N, H * W_, C, _ptr(y), _ptr(ws), ws.numel(), _stream()),
              "global_avgpool_forward")
        ctx.cfg = (x.shape, x.dtype)
        return y

    @staticmethod
    def backward(ctx, g):
        shape, dtype = ctx.cfg
        N, H, W_, C = shape
        g = g.contiguous().float()
        dx = torch.empty(shape, dtype=dtype, device=g.device)
        check(lib.da_global_avgpool_backward(_ptr(g), N, H * W_, C, _ptr(dx), _code(dtype), _stream()), "global_avgpool_backward")
        return dx


def global_avgpool(x):
    return GlobalAvgPoolFunction.apply(x)


class SoftmaxDim0Function(Function):
    """nn.Softmax(dim=1) of NonLocalBlock on [b,q,k] == softmax over the query axis (Q11)."""

    @staticmethod
    def forward(ctx, s):
        _require_cuda(s)
        s = s.contiguous().float()
        T = s.shape[0]
        if s.dim() != 2 or s.shape[1] != T:
            raise RuntimeError("softmax_dim0 expects a square [T,T] matrix")
        p = torch.empty_like(s)
        check(lib.da_softmax_dim0_forward(_ptr(s), T, T, _ptr(p), _stream()), "softmax_dim0_forward")
        ctx.save_for_backward(p)
        return p

    @staticmethod
    def backward(ctx, dp):
        (p,) = ctx.saved_tensors
        dp = dp.contiguous().float()
        ds = torch.empty_like(p)
        check(lib.da_softmax_dim0_backward(_ptr(p), _ptr(dp), p.shape[0], p.shape[0], _ptr(ds), _stream()),
              "softmax_dim0_backward")
        return ds


def softmax_dim0(s):
    return SoftmaxDim0Function.apply(s)


# --------------------------------------------------------------------------------------
# RPN proposal stage: scores, decode, NMS, top max_per_img -- no host round trip
# --------------------------------------------------------------------------------------
def rpn_proposals(cls, reg, base_anchors, stride, img_shape, nms_pre=12000, max_per_img=2000, iou_thr=0.7, min_size=0.0,
                  means=(0., 0., 0., 0.), stds=(1., 1., 1., 1.), wh_ratio_clip=16 / 1000, return_keep=False):
    """One image, one level of RPNHeadDA._get_bboxes_single + _bbox_post_process (rpn_head_da.py:170-303).
    cls [A,H,W] logits, reg [4A,H,W] deltas (the conv outputs as they lie), base_anchors [A,4] ->
    (dets [max_per_img,5] zero padded, count: 0-dim int32 DEVICE tensor[, keep ranks int32 [max_per_img]]).
    The ranking is torch.sort(stable, descending) on the device; everything else is csrc/rpn_proposals.cu."""
    _require_cuda(cls, reg)
    A, H, W = cls.shape
    if reg.shape != (4 * A, H, W):
        raise RuntimeError(f"rpn_proposals: reg {tuple(reg.shape)} does not match cls {tuple(cls.shape)}")
    dev = cls.device
    cls = cls.detach().contiguous().float()
    reg = reg.detach().contiguous().float()
    base = base_anchors.detach().to(device=dev, dtype=torch.float32).contiguous()
    total = A * H * W
    scores = torch.empty((total,), dtype=torch.float32, device=dev)
    check(lib.da_rpn_scores(_ptr(cls), A, H * W, _ptr(scores), _stream()), "rpn_scores")
    top_scores, top_idx = torch.sort(scores, descending=True, stable=True)
    n = int(nms_pre) if 0 < int(nms_pre) < total else total
    top_scores, top_idx = top_scores[:n].contiguous(), top_idx[:n].contiguous()
    dets = torch.empty((int(max_per_img), 5), dtype=torch.float32, device=dev)
    count = torch.empty((), dtype=torch.int32, device=dev)
    keep = torch.empty((int(max_per_img),), dtype=torch.int32, device=dev) if return_keep else None
    ws = workspace(lib.da_rpn_proposals_workspace_bytes(n), dev, "rpn")
    f4 = ctypes.c_float * 4
    check(lib.da_rpn_proposals(_ptr(reg), A, H, W, _ptr(base), float(stride), _ptr(top_idx), _ptr(top_scores), n,
                               f4(*[float(v) for v in means]), f4(*[float(v) for v in stds]), abs(math.log(wh_ratio_clip)),
                               float(img_shape[0]), float(img_shape[1]), float(min_size), float(iou_thr), int(max_per_img),
                               _ptr(dets), _ptr(count), _ptr(keep), _ptr(ws), ws.numel(), _stream()), "rpn_proposals")
    return (dets, count, keep) if return_keep else (dets, count)


def rpn_decoded_boxes(n, device):
    """Decoded boxes [n,4] and validity flags [n] of the last rpn_proposals call on this device (pre-NMS set, rank order)."""
    ws = workspace(lib.da_rpn_proposals_workspace_bytes(n), device, "rpn")
    boxes = torch.empty((n, 4), dtype=torch.float32, device=device)
    valid = torch.empty((n,), dtype=torch.uint8, device=device)
    check(lib.da_rpn_proposals_peek(_ptr(ws), n, _ptr(boxes), _ptr(valid), _stream()), "rpn_proposals_peek")
    return boxes, valid.bool()


# --------------------------------------------------------------------------------------
# blocked NonLocalBlock attention: key blocks, two-pass query-axis softmax, recompute in backward
# --------------------------------------------------------------------------------------
def _gemm_nt(a, b, out, engine):
    """out[M,N] = a[M,K] . b[N,K]^T   (forward form of the 1x1 implicit GEMM: x = a, w = b)."""
    M, K = a.shape
    N = b.shape[0]
    desc = _conv_desc(M, 1, 1, K, N, 1, 1, 1, 0, engine, a.dtype, out.dtype)
    ws = workspace(lib.da_conv_workspace_bytes(ctypes.byref(desc)), a.device, "conv")
    check(lib.da_conv_forward(ctypes.byref(desc), _ptr(a), _ptr(b), None, None, 0, 0.0, 0, _ptr(out), _ptr(ws), ws.numel(),
                              _stream()), "conv_forward")
    return out


def _gemm_nn(a, b, out, engine):
    """out[M,N] = a[M,K] . b[K,N]   (data-gradient form: dz = a, w = b as [Cout=K, Cin=N]; b is read where it lies)."""
    M, K = a.shape
    N = b.shape[1]
    desc = _conv_desc(M, 1, 1, N, K, 1, 1, 1, 0, engine, a.dtype, out.dtype)
    ws = workspace(lib.da_conv_workspace_bytes(ctypes.byref(desc)), a.device, "conv")
    check(lib.da_conv_backward_data(ctypes.byref(desc), _ptr(a), _ptr(b), 1.0, _ptr(out), _ptr(ws), ws.numel(), _stream()),
          "conv_backward_data")
    return out


def _gemm_tn(a, b, out, engine):
    """out[K1,K2] (fp32) = a[M,K1]^T . b[M,K2]   (weight-gradient form: dz = a, x = b)."""
    M, K1 = a.shape
    K2 = b.shape[1]
    desc = _conv_desc(M, 1, 1, K2, K1, 1, 1, 1, 0, engine, a.dtype, a.dtype)
    ws = workspace(lib.da_conv_workspace_bytes(ctypes.byref(desc)), a.device, "conv")
    check(lib.da_conv_backward_weight(ctypes.byref(desc), _ptr(b), _ptr(a), _ptr(out), _ptr(ws), ws.numel(), _stream()),
          "conv_backward_weight")
    return out


class BlockedNonLocalAttention(Function):
    """y[q,:] = sum_k P[q,k] g[k,:],  P = softmax over the QUERY axis of theta.phi^T  (NonLocalBlock,
    mmdet/models/backbones/resnet_da_deep.py:402-445, roi_heads/instance_da.py:150-192; Q11), without the T x T matrix.

    Key columns normalise independently, so the keys are walked in blocks of `block_k`: per block one score GEMM
    [T x block_k] (fp32), the two-pass column softmax of csrc/colsoftmax.cu (statistics kept: 2 floats per key), one
    P.g GEMM accumulated in fp32.  Backward recomputes the block's scores and P from the saved statistics (no second
    statistics pass) and runs the five gradient GEMMs of the block; only theta, phi, g and 2T floats are saved, against
    T^2 floats for the unblocked form (4.3 GB per image at T = 32768, C3 of a 1024x2048 input)."""

    @staticmethod
    def forward(ctx, theta, phi, g, block_k, engine):
        _require_cuda(theta, phi, g)
        theta, phi, g = theta.contiguous(), phi.contiguous(), g.contiguous()
        T, I = theta.shape
        dt, dev = theta.dtype, theta.device
        bk = min(int(block_k), T)
        stats = torch.empty((2 * T,), dtype=torch.float32, device=dev)
        s_buf = torch.empty((T * bk,), dtype=torch.float32, device=dev)
        p_buf = torch.empty((T * bk,), dtype=dt, device=dev)
        y = torch.empty((T, I), dtype=torch.float32, device=dev)
        yb = torch.empty((T, I), dtype=torch.float32, device=dev) if T > bk else None
        ws = workspace(lib.da_colsoftmax_workspace_bytes(T, bk), dev, "colsoftmax")
        for k0 in range(0, T, bk):
            tk = min(bk, T - k0)
            s = _gemm_nt(theta, phi[k0:k0 + tk], s_buf[:T * tk].view(T, tk), engine)
            p = p_buf[:T * tk].view(T, tk)
            check(lib.da_colsoftmax_forward(_ptr(s), T, tk, tk, _ptr(p), _code(dt), _ptr(stats[2 * k0:]), 0, _ptr(ws), ws.numel(),
                                            _stream()), "colsoftmax_forward")
            if k0 == 0:
                _gemm_nn(p, g[k0:k0 + tk], y, engine)
            else:
                y += _gemm_nn(p, g[k0:k0 + tk], yb, engine)
        ctx.save_for_backward(theta, phi, g, stats)
        ctx.cfg = (bk, engine)
        return y.to(dt)

    @staticmethod
    def backward(ctx, dy):
        theta, phi, g, stats = ctx.saved_tensors
        bk, engine = ctx.cfg
        T, I = theta.shape
        dt, dev = theta.dtype, theta.device
        dy = cast(dy.contiguous(), dt)
        s_buf = torch.empty((T * bk,), dtype=torch.float32, device=dev)
        dp_buf = torch.empty((T * bk,), dtype=torch.float32, device=dev)
        p_buf = torch.empty((T * bk,), dtype=dt, device=dev)
        ds_buf = torch.empty((T * bk,), dtype=dt, device=dev)
        dtheta = torch.empty((T, I), dtype=torch.float32, device=dev)
        dtb = torch.empty((T, I), dtype=torch.float32, device=dev) if T > bk else None
        dphi = torch.empty((T, I), dtype=torch.float32, device=dev)
        dg = torch.empty((T, I), dtype=torch.float32, device=dev)
        ws = workspace(lib.da_colsoftmax_workspace_bytes(T, bk), dev, "colsoftmax")
        for k0 in range(0, T, bk):
            tk = min(bk, T - k0)
            phi_b, g_b = phi[k0:k0 + tk], g[k0:k0 + tk]
            s = _gemm_nt(theta, phi_b, s_buf[:T * tk].view(T, tk), engine)
            p = p_buf[:T * tk].view(T, tk)
            check(lib.da_colsoftmax_forward(_ptr(s), T, tk, tk, _ptr(p), _code(dt), _ptr(stats[2 * k0:]), 1, None, 0, _stream()),
                  "colsoftmax_forward")
            dp = _gemm_nt(dy, g_b, dp_buf[:T * tk].view(T, tk), engine)
            _gemm_tn(p, dy, dg[k0:k0 + tk], engine)                       # dg_b = P_b^T . dy
            ds = ds_buf[:T * tk].view(T, tk)
            check(lib.da_colsoftmax_backward(_ptr(p), _code(dt), _ptr(dp), T, tk, tk, _ptr(ds), _code(dt), _ptr(ws), ws.numel(),
                                             _stream()), "colsoftmax_backward")
            if k0 == 0:
                _gemm_nn(ds, phi_b, dtheta, engine)                       # dtheta += dS_b . phi_b
            else:
                dtheta += _gemm_nn(ds, phi_b, dtb, engine)
            _gemm_tn(ds, theta, dphi[k0:k0 + tk], engine)                 # dphi_b = dS_b^T . theta
        return dtheta.to(dt), dphi.to(dt), dg.to(dt), None, None


def nonlocal_attention_blocked(theta, phi, g, block_k=4096, engine=None):
    """theta, phi, g: [T,I] token projections in the activation dtype -> y [T,I].  T and I must be multiples of 8 (tile
    granularity of the tensor-core engines; the CUDA-core engine takes any size)."""
    engine = engine or get_engine()
    T, I = theta.shape
    if engine != "simt_f32" and (T % 8 or I % 8 or int(block_k) % 8):
        raise RuntimeError(f"nonlocal_attention_blocked: T={T}, I={I}, block_k={block_k} must be multiples of 8 on engine {engine}")
    return BlockedNonLocalAttention.apply(theta, phi, g, int(block_k), engine)


# --------------------------------------------------------------------------------------
# instance-level domain classifier + CE as one persistent kernel (csrc/chain.cu)
# --------------------------------------------------------------------------------------
def pack_projection(weights):
    """The three 1x1 projection weights of a NonLocalBlock (theta | phi | g, each [I,C,1,1]) as ONE [3I,C] buffer, and their
    bf16 operand copies as one [3I,C] buffer: the chain kernel runs the three projections (and their data / weight gradients)
    as single GEMMs.  The parameters keep their identity, names and shapes; only their storage is re-seated (once; again after
    a .to()/.cuda() that separated them).  Returns (fp32 [3I,C] view of the masters, bf16 [3I,C] operand copy)."""
    I_, C = weights[0].shape[0], weights[0].shape[1]
    n = I_ * C
    base = weights[0].data_ptr()
    packed = all(w.dtype == torch.float32 and w.data_ptr() == base + 4 * n * i and _dense_memory(w) for i, w in enumerate(weights))
    hit = getattr(weights[0], "_da_pack", None)
    if not packed or hit is None or hit[0].data_ptr() != base:
        with torch.no_grad():
            buf = torch.empty((3 * I_, C), dtype=torch.float32, device=weights[0].device)
            for i, w in enumerate(weights):
                buf[i * I_:(i + 1) * I_].copy_(w.detach().reshape(I_, C))
                w.data = buf[i * I_:(i + 1) * I_].view(I_, C, 1, 1)
            sh = torch.empty((3 * I_, C), dtype=torch.bfloat16, device=buf.device)
            check(lib.da_cast(_ptr(buf), _lib.DA_F32, _ptr(sh), _lib.DA_BF16, buf.numel(), _stream()), "cast")
            for i, w in enumerate(weights):
                w._da_shadow = (w._version, sh[i * I_:(i + 1) * I_].view(I_, C, 1, 1))
            weights[0]._da_pack = (buf, sh)
        hit = weights[0]._da_pack
    buf, sh = hit
    for i, w in enumerate(weights):      # a torch in-place update of a master invalidates its operand copy: refresh
        cur = getattr(w, "_da_shadow", None)
        if cur is None or cur[0] != w._version or cur[1].data_ptr() != sh.data_ptr() + 2 * n * i:
            with torch.no_grad():
                view = sh[i * I_:(i + 1) * I_].view(I_, C, 1, 1)
                check(lib.da_cast(_ptr(w.detach()), _lib.DA_F32, _ptr(view), _lib.DA_BF16, n, _stream()), "cast")
                w._da_shadow = (w._version, view)
    return buf, sh


class InstanceHeadChainFunction(Function):
    """(loss, pred) = CE-on-sigmoid( FC stack ( [NonLocalBlock]( GRL(x) ) ) ) in one kernel; backward in one kernel."""

    @staticmethod
    def forward(ctx, x, labels, nlb_w, w_mask, w1, b1, w2, b2, w3, b3, ops, cfg, xin=None, w0=None, b0=None, b_in=None, op0=None):
        # nlb_w: packed fp32 [3I,C] view of the projection masters (or None); ops: bf16 operand copies in the same order
        # feeding layer (x is None): x = relu(xin w0^T + b0) inside the kernel; b_in = bias of the layer that produced xin
        # (a post-ReLU activation) when that layer runs with preact_grad=True: its gradient is returned through b_in
        pre = xin is not None
        if pre:
            _require_cuda(xin)
            xin = xin.contiguous()
            R, C0 = xin.shape
            C = w0.shape[0]
            x = torch.empty((R, C), dtype=torch.bfloat16, device=xin.device)
        else:
            _require_cuda(x)
            R, C = x.shape
            C0 = 0
        drop_p, seed1, seed2, grl = cfg
        nlb = nlb_w is not None
        I_ = nlb_w.shape[0] // 3 if nlb else 0
        H1, H2 = w1.shape[0], w2.shape[0]
        dev = x.device
        x = x.contiguous()
        lab = _as_i32(labels, dev)
        bf, f32 = torch.bfloat16, torch.float32
        ldp = (R + 7) // 8 * 8
        sv = dict(h1=torch.empty((R, H1), dtype=bf, device=dev), h2=torch.empty((R, H2), dtype=bf, device=dev),
                  z=torch.empty((R, 2), dtype=f32, device=dev))
        if nlb:
            sv.update(proj=torch.empty((R, 3 * I_), dtype=bf, device=dev), attn=torch.empty((R, ldp), dtype=bf, device=dev),
                      y=torch.empty((R, I_), dtype=bf, device=dev), t=torch.empty((R, C), dtype=bf, device=dev))
        pred = torch.empty((R, 2), dtype=f32, device=dev)
        loss = torch.empty((), dtype=f32, device=dev)
        op_proj, op_mask, op1, op2, op3 = ops
        b1c, b2c, b3c = (b.detach().float().contiguous() for b in (b1, b2, b3))
        b0c = b0.detach().float().contiguous() if pre else None
        desc = _lib.InstanceFcDesc(R, C, I_, H1, H2, int(nlb), float(drop_p), int(seed1), int(seed2), float(grl), C0, int(b_in is not None))
        g = lambda t_: None if t_ is None else t_.data_ptr()
        ten = _lib.InstanceFcTensors(g(x), g(op_proj), g(op_mask), g(op1), g(b1c), g(op2), g(b2c), g(op3), g(b3c), g(lab), g(sv.get("proj")),
                                     g(sv.get("attn")), g(sv.get("y")), g(sv.get("t")), g(sv["h1"]), g(sv["h2"]), g(sv["z"]), g(pred), g(loss),
                                     g(xin), g(op0), g(b0c))
        ws = workspace(lib.da_instance_fc_workspace_bytes(R), dev, "chain")
        check(lib.da_instance_fc_forward(ctypes.byref(desc), ctypes.byref(ten), _ptr(ws), ws.numel(), _stream()), "instance_fc_forward")
        ctx.desc, ctx.ten, ctx.keep = desc, ten, (x, lab, sv, ops, b1c, b2c, b3c, pred, xin, op0, b0c)
        ctx.shapes = (None if not nlb else nlb_w.shape, None if w_mask is None else w_mask.shape, w1.shape, w2.shape, w3.shape,
                      None if not pre else w0.shape)
        return loss, pred

    @staticmethod
    def backward(ctx, g_loss, g_pred):
        x, lab, sv, ops, b1c, b2c, b3c, pred, xin, op0, b0c = ctx.keep
        desc = ctx.desc
        R, C, I_, H1, H2, nlb = desc.R, desc.C, desc.I, desc.H1, desc.H2, bool(desc.nlb)
        C0, gate_in = desc.C0, bool(desc.gate_in)
        dev = x.device
        bf, f32 = torch.bfloat16, torch.float32
        gl = None if g_loss is None else g_loss.contiguous().float()
        gp = None if g_pred is None else g_pred.contiguous().float()
        dx = torch.empty((R, C), dtype=bf, device=dev)
        dw1, db1 = torch.empty((H1, C), dtype=f32, device=dev), torch.empty((H1,), dtype=f32, device=dev)
        dw2, db2 = torch.empty((H2, H1), dtype=f32, device=dev), torch.empty((H2,), dtype=f32, device=dev)
        dw3, db3 = torch.empty((2, H2), dtype=f32, device=dev), torch.empty((2,), dtype=f32, device=dev)
        dz2, dz1 = torch.empty((R, H2), dtype=bf, device=dev), torch.empty((R, H1), dtype=bf, device=dev)
        dwp = dwm = dt = dy = dproj = None
        if nlb:
            dwp, dwm = torch.empty((3 * I_, C), dtype=f32, device=dev), torch.empty((C, I_), dtype=f32, device=dev)
            dt, dy = torch.empty((R, C), dtype=bf, device=dev), torch.empty((R, I_), dtype=bf, device=dev)
            dproj = torch.empty((R, 3 * I_), dtype=bf, device=dev)
        dxin = dw0 = db0 = db_in = None
        if C0:
            dxin = torch.empty((R, C0), dtype=bf, device=dev)
            dw0, db0 = torch.empty((C, C0), dtype=f32, device=dev), torch.empty((C,), dtype=f32, device=dev)
            if gate_in:
                db_in = torch.empty((C0,), dtype=f32, device=dev)
        g = lambda t_: None if t_ is None else t_.data_ptr()
        gr = _lib.InstanceFcGrads(g(gl), 1.0, g(gp), g(dx), g(dwp), g(dwm), g(dw1), g(db1), g(dw2), g(db2), g(dw3), g(db3), g(dz2), g(dz1),
                                  g(dt), g(dy), g(dproj), g(dxin), g(dw0), g(db0), g(db_in))
        ws = workspace(lib.da_instance_fc_workspace_bytes(R), dev, "chain")
        check(lib.da_instance_fc_backward(ctypes.byref(desc), ctypes.byref(ctx.ten), ctypes.byref(gr), _ptr(ws), ws.numel(), _stream()),
              "instance_fc_backward")
        s_proj, s_mask, s1, s2, s3, s0 = ctx.shapes
        return (None if C0 else dx, None, dwp, None if dwm is None else dwm.view(s_mask), dw1.view(s1), db1, dw2.view(s2), db2,
                dw3.view(s3), db3, None, None, dxin, None if dw0 is None else dw0.view(s0), db0, db_in, None)


class _SplitPacked(Function):
    """Identity bridge between the three projection parameters and their packed [3I,C] view: backward hands each parameter
    its slice of the packed gradient (views of one buffer, no copies)."""

    @staticmethod
    def forward(ctx, packed_view, w_theta, w_phi, w_g):
        ctx.shapes = (w_theta.shape, w_phi.shape, w_g.shape)
        return packed_view.detach()

    @staticmethod
    def backward(ctx, d):
        I_ = d.shape[0] // 3
        parts = [d[i * I_:(i + 1) * I_].view(s) for i, s in enumerate(ctx.shapes)]
        return (None, *parts)


def instance_head_chain(x, labels, nlb_weights, w_mask, fcs, drop_p, seeds, grl, pre=None):
    """x [R,C] bf16; nlb_weights: (theta, phi, g) conv weights or None; w_mask: conv_mask weight or None;
    fcs: ((w1,b1),(w2,b2),(w3,b3)).  -> (loss = mean CE(sigmoid(fc3), labels), pred = sigmoid(fc3) [R,2]).
    pre = (xin [R,C0] bf16, w0 [C,C0], b0 [C], b_in or None): the layer that PRODUCES the features runs inside the kernel
    (x = relu(xin w0^T + b0); pass x=None); with b_in (the bias of the layer that produced xin, itself called with
    dense_layer(..., preact_grad=True)) the kernel also applies that layer's ReLU derivative and returns its bias gradient."""
    (w1, b1), (w2, b2), (w3, b3) = fcs
    if pre is not None:
        xin, w0, b0, b_in = pre
        if x is not None or xin.dim() != 2 or w0.shape[1] != xin.shape[1] or xin.shape[1] % 64 or (b_in is not None and b_in.shape[0] != xin.shape[1]):
            raise RuntimeError("instance_head_chain: bad feeding layer (x must be None, xin [R,C0], w0 [C,C0], C0 % 64 == 0)")
        C = w0.shape[0]
    else:
        C = x.shape[1]
    if nlb_weights is not None and (any(w.shape[1] != C for w in nlb_weights) or w_mask.shape[0] != C or w1.shape[1] != C):
        raise RuntimeError(f"instance_head_chain: features are {C} wide, the NonLocalBlock / fc1 expect {nlb_weights[0].shape[1]} / {w1.shape[1]}")
    if w1.shape[1] != C or w2.shape[1] != w1.shape[0] or w3.shape[1] != w2.shape[0] or w3.shape[0] != 2:
        raise RuntimeError("instance_head_chain: FC shapes do not chain (C -> H1 -> H2 -> 2)")
    if nlb_weights is not None:
        buf, sh = pack_projection(list(nlb_weights))
        nlb_w = _SplitPacked.apply(buf, *nlb_weights)
        op_proj, op_mask = sh, bf16_shadow(w_mask)
    else:
        nlb_w, op_proj, op_mask = None, None, None
    ops = (op_proj, op_mask, bf16_shadow(w1), bf16_shadow(w2), bf16_shadow(w3))
    cfg = (drop_p, seeds[0], seeds[1], grl)
    if pre is not None:
        return InstanceHeadChainFunction.apply(None, labels, nlb_w, w_mask, w1, b1, w2, b2, w3, b3, ops, cfg, xin, w0, b0, b_in,
                                               bf16_shadow(w0))
    return InstanceHeadChainFunction.apply(x, labels, nlb_w, w_mask, w1, b1, w2, b2, w3, b3, ops, cfg, None, None, None, None, None)


# --------------------------------------------------------------------------------------
# pixel-level domain classifier tail: producing conv -> 1-channel conv -> per-pixel loss -> mean (csrc/pixel_tail.cu)
# --------------------------------------------------------------------------------------
class ConvPixelLossFunction(Function):
    """(loss, logits) = tail(act(conv(x, w) * scale + shift)): the last two layers of a pixel-level domain classifier and its
    loss (da_grl_conv_loss_forward / _backward).  `grl` multiplies the gradient into x (first layer of a head) -- pass 1.0
    for an inner layer."""

    @staticmethod
    def forward(ctx, x, w, scale, shift, w_tail, b_tail, domain, cfg, shadow):
        stride, pad, relu, drop_p, seed, engine, grl, tail_relu, mode, gamma, alpha = cfg
        _require_cuda(x, w)
        N, H, W_, Cin = x.shape
        x = x.contiguous()
        if shadow is not None:
            wv = _w_ohwi(shadow)
        else:
            wv = _w_ohwi(w.detach())
            if wv.dtype != x.dtype:
                wv = cast(wv, x.dtype)
        Cout, KH, KW, _ = wv.shape
        OH, OW = (H + 2 * pad - KH) // stride + 1, (W_ + 2 * pad - KW) // stride + 1
        dev = x.device
        y = torch.empty((N, OH, OW, Cout), dtype=x.dtype, device=dev)
        logits = torch.empty((N, OH, OW), dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        desc = _conv_desc(N, H, W_, Cin, Cout, KH, KW, stride, pad, engine, x.dtype, x.dtype)
        ws = workspace(lib.da_grl_conv_loss_workspace_bytes(ctypes.byref(desc)), dev, "conv_tail")
        sc = None if scale is None else scale.detach().float().contiguous()
        sh = None if shift is None else shift.detach().float().contiguous()
        wt = w_tail.detach().reshape(-1).float().contiguous()
        bt = None if b_tail is None else b_tail.detach().reshape(-1).float().contiguous()
        dom = _as_i32(domain, dev)
        tail = _lib.PixelTail(wt.data_ptr(), None if bt is None else bt.data_ptr(), int(tail_relu), _lib.PIXEL_LOSS_MODES[mode],
                              float(gamma), float(alpha), dom.data_ptr())
        check(lib.da_grl_conv_loss_forward(ctypes.byref(desc), _ptr(x), _ptr(wv), _ptr(sc), _ptr(sh), int(relu), float(drop_p), int(seed),
                                           _ptr(y), ctypes.byref(tail), _ptr(logits), _ptr(loss), _ptr(ws), ws.numel(), _stream()),
              "grl_conv_loss_forward")
        ctx.save_for_backward(x, wv, y, sc, sh, wt, bt, dom, logits)
        ctx.cfg, ctx.desc = cfg, desc
        ctx.meta = (w.shape, w.dtype, w_tail.shape, b_tail is not None)
        return loss, logits

    @staticmethod
    def backward(ctx, g_loss, g_logits):
        x, wv, y, sc, sh, wt, bt, dom, logits = ctx.saved_tensors
        stride, pad, relu, drop_p, seed, engine, grl, tail_relu, mode, gamma, alpha = ctx.cfg
        wshape, wdtype, wt_shape, has_bt = ctx.meta
        dev = x.device
        Cout, KH, KW, Cin = wv.shape
        gl = None if g_loss is None else g_loss.contiguous().float()
        gq = None if g_logits is None else g_logits.contiguous().float()
        need_x, need_w, need_scale, need_shift = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2], ctx.needs_input_grad[3]
        dx = torch.empty_like(x) if need_x else None
        dwv = torch.empty((Cout, KH, KW, Cin), dtype=torch.float32, device=dev) if need_w else None
        dshift = torch.empty((Cout,), dtype=torch.float32, device=dev)
        dvdot = torch.empty((Cout,), dtype=torch.float32, device=dev)
        dwt = torch.empty((Cout,), dtype=torch.float32, device=dev)
        dbt = torch.empty((1,), dtype=torch.float32, device=dev)
        dz = torch.empty_like(y)
        tail = _lib.PixelTail(wt.data_ptr(), None if bt is None else bt.data_ptr(), int(tail_relu), _lib.PIXEL_LOSS_MODES[mode],
                              float(gamma), float(alpha), dom.data_ptr())
        ws = workspace(lib.da_grl_conv_loss_workspace_bytes(ctypes.byref(ctx.desc)), dev, "conv_tail")
        check(lib.da_grl_conv_loss_backward(ctypes.byref(ctx.desc), _ptr(x), _ptr(wv), _ptr(sc), int(relu), float(drop_p), _ptr(y),
                                            ctypes.byref(tail), _ptr(logits), _ptr(gl), 1.0, _ptr(gq), float(grl), _ptr(dx), _ptr(dwv),
                                            _ptr(dshift), _ptr(dvdot), _ptr(dwt), _ptr(dbt), _ptr(dz), _ptr(ws), ws.numel(), _stream()),
              "grl_conv_loss_backward")
        dw = None
        if need_w:
            dw = dwv.view(wshape) if len(wshape) == 2 else dwv.permute(0, 3, 1, 2)
            if wdtype != torch.float32:
                dw = dw.to(wdtype)
            elif WGRAD_HOOK is not None:
                WGRAD_HOOK(dwv)
        dscale = None
        if need_scale:
            t = sh if sh is not None else torch.zeros_like(dshift)
            safe = torch.where(sc == 0, torch.ones_like(sc), sc)
            dscale = torch.where(sc == 0, torch.zeros_like(sc), (dvdot - t * dshift) / safe)
        return (dx, dw, dscale, dshift if need_shift else None, dwt.view(wt_shape), dbt if has_bt else None, None, None, None)


def conv_pixel_loss(x, w, scale, shift, w_tail, b_tail, domain, stride=1, pad=0, relu=True, drop_p=0.0, seed=0, engine=None, grl=1.0,
                    tail_relu=False, mode="daf_sq_batch", gamma=2.0, alpha=0.25):
    """x [N,H,W,Cin] NHWC; w the producing conv's weight; (scale, shift): bias / folded BN; w_tail [1,Cout,1,1], b_tail [1] or None;
    domain [N].  -> (loss scalar, logits [N,OH,OW] fp32).  mode: 'daf_sq_batch' (L1), 'daf_sq_image' (L2), 'bce', 'focal'."""
    engine = engine or get_engine()
    if engine != "simt_f32" and not (_umma_ok(x, w) and stride <= 2):
        engine = "simt_f32"
    if mode not in _lib.PIXEL_LOSS_MODES:
        raise ValueError(f"unknown pixel loss mode {mode!r}; choose from {sorted(_lib.PIXEL_LOSS_MODES)}")
    shadow = None
    if w.dtype == torch.float32 and x.dtype == torch.bfloat16 and w.is_leaf and _dense_memory(w):
        shadow = bf16_shadow(w)
    cfg = (stride, pad, int(relu), float(drop_p), int(seed), engine, float(grl), int(tail_relu), mode, float(gamma), float(alpha))
    return ConvPixelLossFunction.apply(x, w, scale, shift, w_tail, b_tail, domain, cfg, shadow)


# --------------------------------------------------------------------------------------
# W1: lambda-weighted loss entries + their total in one launch
# --------------------------------------------------------------------------------------
class WeightedLossesFunction(Function):
    @staticmethod
    def forward(ctx, weights, *losses):
        n = len(losses)
        dev = losses[0].device
        ls = [l.detach().float().contiguous() for l in losses]
        ptrs = (ctypes.c_void_p * n)(*[l.data_ptr() for l in ls])
        w = (ctypes.c_float * n)(*[float(x) for x in weights])
        scaled = torch.empty((n,), dtype=torch.float32, device=dev)
        total = torch.empty((), dtype=torch.float32, device=dev)
        check(lib.da_weighted_sum_forward(ptrs, w, n, _ptr(scaled), _ptr(total), _stream()), "weighted_sum_forward")
        ctx.weights, ctx.n = tuple(float(x) for x in weights), n
        return scaled, total

    @staticmethod
    def backward(ctx, g_scaled, g_total):
        n = ctx.n
        dev = (g_total if g_total is not None else g_scaled).device
        w = (ctypes.c_float * n)(*ctx.weights)
        d = torch.empty((n,), dtype=torch.float32, device=dev)
        gs = None if g_scaled is None else g_scaled.contiguous().float()
        gt = None if g_total is None else g_total.contiguous().float()
        check(lib.da_weighted_sum_backward(w, n, _ptr(gt), _ptr(gs), _ptr(d), _stream()), "weighted_sum_backward")
        return (None, *[d[i] for i in range(n)])


def weighted_losses(losses, weights):
    """-> (scaled [n] with scaled[i] = weights[i] * losses[i], total = scaled.sum()), one kernel forward, one backward."""
    _require_cuda(*losses)
    return WeightedLossesFunction.apply(tuple(weights), *losses)
