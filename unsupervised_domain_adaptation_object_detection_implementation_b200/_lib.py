"""ctypes binding of libda_b200.so (the C-ABI declared in include/da_b200.h).

The library is the product: there is no CPU or PyTorch fallback.  Importing this module
without a built library raises immediately ("fail loudly"), and every wrapper raises
RuntimeError with da_last_error() when a call returns non-zero -- the same error surface
mmcv's ext_module gives the reference (AT_CUDA_CHECK -> RuntimeError).
"""
import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t,
                    c_uint64, c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libda_b200.so")

# enums of include/da_b200.h
DA_F32, DA_BF16 = 0, 1
ROI_OUT_RCHW, ROI_OUT_RHWC = 0, 1
ENGINE_SIMT_F32, ENGINE_UMMA_BF16, ENGINE_UMMA_BF16X3, ENGINE_UMMA_BF16X6 = 0, 1, 2, 3
ENGINES = {"simt_f32": ENGINE_SIMT_F32, "umma_bf16": ENGINE_UMMA_BF16, "umma_bf16x3": ENGINE_UMMA_BF16X3,
           "umma_bf16x6": ENGINE_UMMA_BF16X6}


class ConvDesc(Structure):
    _fields_ = [("N", c_int), ("H", c_int), ("W", c_int), ("Cin", c_int),
                ("Cout", c_int), ("KH", c_int), ("KW", c_int),
                ("stride", c_int), ("pad", c_int),
                ("engine", c_int), ("x_dtype", c_int), ("y_dtype", c_int)]


DA_MAX_PEERS, DA_PEER_HANDLE_BYTES, DA_PEER_FLAG_INTS = 8, 64, 16
DA_PEER_PUBLISH_STORES, DA_PEER_PUBLISH_BY_CALLER = 0, 1


class PeerSgdArgs(Structure):
    """da_peer_sgd_args (include/da_b200.h)."""
    _fields_ = [("w", c_void_p), ("momentum_shard", c_void_p),
                ("grad", c_void_p * DA_MAX_PEERS), ("w_bf16", c_void_p * DA_MAX_PEERS),
                ("w_f32", c_void_p * DA_MAX_PEERS), ("flags", c_void_p * DA_MAX_PEERS),
                ("local_state", c_void_p), ("n", c_int64), ("world", c_int32), ("rank", c_int32)]


class SgdFuse(Structure):
    """da_sgd_fuse (include/da_b200.h)."""
    _fields_ = [("w", c_void_p), ("momentum_buf", c_void_p), ("w_bf16", c_void_p),
                ("lr", c_float), ("momentum", c_float), ("weight_decay", c_float), ("first_step", c_int32)]


class InstanceFcDesc(Structure):
    """da_instance_fc_desc (include/da_b200.h)."""
    _fields_ = [("R", c_int), ("C", c_int), ("I", c_int), ("H1", c_int), ("H2", c_int), ("nlb", c_int), ("drop_p", c_float),
                ("seed1", c_uint64), ("seed2", c_uint64), ("grl", c_float), ("C0", c_int), ("gate_in", c_int)]


class InstanceFcTensors(Structure):
    """da_instance_fc_tensors."""
    _fields_ = [(n, c_void_p) for n in ("x", "w_proj", "w_mask", "w1", "b1", "w2", "b2", "w3", "b3", "labels", "proj", "attn", "y", "t",
                                        "h1", "h2", "z", "pred", "loss", "xin", "w0", "b0")]


class InstanceFcGrads(Structure):
    """da_instance_fc_grads."""
    _fields_ = [("grad_loss", c_void_p), ("loss_scale", c_float), ("grad_pred", c_void_p), ("dx", c_void_p), ("dw_proj", c_void_p),
                ("dw_mask", c_void_p), ("dw1", c_void_p), ("db1", c_void_p), ("dw2", c_void_p), ("db2", c_void_p), ("dw3", c_void_p),
                ("db3", c_void_p), ("dz2", c_void_p), ("dz1", c_void_p), ("dt", c_void_p), ("dy", c_void_p), ("dproj", c_void_p),
                ("dxin", c_void_p), ("dw0", c_void_p), ("db0", c_void_p), ("db_in", c_void_p)]


class PixelTail(Structure):
    """da_pixel_tail."""
    _fields_ = [("w", c_void_p), ("bias", c_void_p), ("relu", c_int), ("mode", c_int), ("gamma", c_float), ("alpha", c_float),
                ("domain", c_void_p)]


PIXEL_LOSS_MODES = {"daf_sq_batch": 0, "daf_sq_image": 1, "bce": 2, "focal": 3}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C unsupervised_domain_adaptation_object_detection_implementation_b200/csrc`. "
            "There is no CPU fallback for the DA hot path.")
    return ctypes.CDLL(LIB_PATH)


lib = _load()

P, I, F, S, L, U64 = c_void_p, c_int, c_float, c_size_t, c_int64, c_uint64
CD = POINTER(ConvDesc)

# name -> (restype, argtypes); every symbol include/da_b200.h declares
SIGNATURES = {
    "da_version": (I, []),
    "da_last_error": (c_char_p, []),
    "da_launch_count": (L, []),
    "da_launch_count_reset": (None, []),
    "da_set_option": (I, [c_char_p, L]),
    "da_grl_backward": (I, [P, P, I, L, F, P]),
    "da_nchw_to_nhwc": (I, [P, I, P, I, I, I, I, I, P]),
    "da_nhwc_to_nchw": (I, [P, I, P, I, I, I, I, I, P]),
    "da_split_bf16": (I, [P, P, P, L, P]),
    "da_cast": (I, [P, I, P, I, L, P]),
    "da_sgd_step": (I, [P, P, P, L, F, F, F, I, P, P]),
    "da_set_sm_limit": (I, [I]),
    "da_sgd_step_multi": (I, [P, I, P, I, F, F, F, P]),
    "da_sgd_step_peer": (I, [P, F, F, F, I, I, I, P]),
    "da_peer_copy": (I, [P, P, S, P]),
    "da_peer_publish_done": (I, [P, P]),
    "da_peer_signal_done": (I, [P, P]),
    "da_peer_wait_done": (I, [P, P]),
    "da_peer_alloc": (I, [S, P]),
    "da_peer_free": (I, [P]),
    "da_peer_export": (I, [P, P]),
    "da_peer_open": (I, [P, P]),
    "da_peer_close": (I, [P]),
    "da_roi_align_workspace_bytes": (S, [I, I, I]),
    "da_roi_align_forward": (I, [P, I, I, I, I, I, P, I, I, I, F, I, I, P, I, I, P, P, S, P]),
    "da_roi_align_backward": (I, [P, I, I, P, I, I, I, F, I, I, P, I, I, I, I, I, P, S, P]),
    "da_roi_align_backward_prepared": (I, [P, I, I, P, I, I, I, F, I, I, P, I, I, I, I, I, P, S, P]),
    "da_map_roi_levels": (I, [P, I, I, F, P, P]),
    "da_pixel_loss_workspace_bytes": (S, [I, L]),
    "da_pixel_domain_loss_forward": (I, [P, I, L, P, I, P, P, S, P]),
    "da_pixel_domain_loss_backward": (I, [P, I, L, P, I, P, F, P, P]),
    "da_ce2_forward": (I, [P, P, I, I, P, P, P]),
    "da_ce2_backward": (I, [P, P, I, I, P, F, P, P, P]),
    "da_focal2_forward": (I, [P, P, I, F, F, P, P]),
    "da_focal2_backward": (I, [P, P, I, F, F, P, F, P, P]),
    "da_consistency_forward": (I, [P, L, P, P, I, P, P, P]),
    "da_consistency_backward": (I, [P, L, P, P, I, P, P, F, P, P, P]),
    "da_pixel_head_forward": (I, [P, I, L, I, P, P, I, P, P]),
    "da_pixel_head_workspace_bytes": (S, [L, I]),
    "da_pixel_head_backward": (I, [P, I, L, I, P, P, P, P, I, P, P, P, S, P]),
    "da_conv_workspace_bytes": (S, [CD]),
    "da_conv_forward": (I, [CD, P, P, P, P, I, F, U64, P, P, S, P]),
    "da_conv_act_backward": (I, [CD, P, P, P, I, F, U64, P, P, P, P, S, P]),
    "da_conv_backward_data": (I, [CD, P, P, F, P, P, S, P]),
    "da_conv_backward_weight": (I, [CD, P, P, P, P, S, P]),
    "da_conv_backward_weight_sgd": (I, [CD, P, P, P, P, S, P]),
    "da_dropout_mask": (I, [U64, L, F, P, P]),
    "da_set_dropout_counter": (I, [P]),
    "da_global_avgpool_workspace_bytes": (S, [I, I]),
    "da_global_avgpool_forward": (I, [P, I, I, I, I, P, P, S, P]),
    "da_global_avgpool_backward": (I, [P, I, I, I, P, I, P]),
    "da_softmax_dim0_forward": (I, [P, I, I, P, P]),
    "da_softmax_dim0_backward": (I, [P, P, I, I, P, P]),
    "da_rpn_scores": (I, [P, I, I, P, P]),
    "da_rpn_proposals_workspace_bytes": (S, [I]),
    "da_rpn_proposals": (I, [P, I, I, I, P, F, P, P, I, POINTER(c_float), POINTER(c_float), F, F, F, F, F, I, P, P, P, P, S, P]),
    "da_rpn_proposals_peek": (I, [P, I, P, P, P]),
    "da_colsoftmax_workspace_bytes": (S, [I, I]),
    "da_colsoftmax_forward": (I, [P, I, I, I, P, I, P, I, P, S, P]),
    "da_colsoftmax_backward": (I, [P, I, P, I, I, I, P, I, P, S, P]),
    "da_weighted_sum_forward": (I, [POINTER(c_void_p), POINTER(c_float), I, P, P, P]),
    "da_weighted_sum_backward": (I, [POINTER(c_float), I, P, P, P, P]),
    "da_grl_conv_loss_workspace_bytes": (S, [CD]),
    "da_pixel_tail_forward": (I, [CD, P, POINTER(PixelTail), P, P, P, S, P]),
    "da_pixel_tail_backward": (I, [CD, P, POINTER(PixelTail), P, P, F, P, P, I, F, P, P, P, P, P, P, S, P]),
    "da_grl_conv_loss_forward": (I, [CD, P, P, P, P, I, F, U64, P, POINTER(PixelTail), P, P, P, S, P]),
    "da_grl_conv_loss_backward": (I, [CD, P, P, P, I, F, P, POINTER(PixelTail), P, P, F, P, F, P, P, P, P, P, P, P, P, S, P]),
    "da_instance_fc_workspace_bytes": (S, [I]),
    "da_instance_fc_forward": (I, [POINTER(InstanceFcDesc), POINTER(InstanceFcTensors), P, S, P]),
    "da_instance_fc_backward": (I, [POINTER(InstanceFcDesc), POINTER(InstanceFcTensors), POINTER(InstanceFcGrads), P, S, P]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here == header/library mismatch: fail loudly
    _fn.restype = _res
    _fn.argtypes = _args


def last_error():
    return lib.da_last_error().decode("utf-8", "replace")


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError(f"libda_b200 {what} failed (code {rc}): {last_error()}")


def launch_count():
    return int(lib.da_launch_count())


def reset_launch_count():
    lib.da_launch_count_reset()
