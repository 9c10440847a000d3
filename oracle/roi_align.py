"""ORACLE — test infrastructure only; never imported by the product path.

ctypes front-end of oracle/roi_align_ref.c (CPU restatement of mmcv RoIAlign avg/aligned as
called at /root/reference/mmdet/models/roi_heads/roi_extractors/single_level_roi_extractor.py:79).
Parity status: PINNED against torch.ops.torchvision.roi_align (C++ CPU op) in
tests/test_oracle_cpu.py and against tests/golden/roi_align_*.pt.
"""
import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libda_oracle.so")


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)
    return _SO


def _lib():
    if not os.path.exists(_SO):
        build()
    lib = ctypes.CDLL(_SO)
    fp, ip = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int32)
    lib.roi_align_forward_ref_range.argtypes = [fp] + [ctypes.c_int] * 4 + [fp] + [ctypes.c_int] * 4 + \
        [ctypes.c_float, ctypes.c_int, ctypes.c_int, fp, ip, ip]
    lib.roi_align_backward_ref.argtypes = [fp, fp] + [ctypes.c_int] * 3 + [ctypes.c_float, ctypes.c_int, ctypes.c_int,
                                                                          ctypes.POINTER(ctypes.c_double)] + [ctypes.c_int] * 4
    lib.roi_align_backward_ref_range.argtypes = [fp, fp] + [ctypes.c_int] * 3 + [ctypes.c_float, ctypes.c_int, ctypes.c_int,
                                                                                ctypes.POINTER(ctypes.c_double)] + [ctypes.c_int] * 6
    lib.roi_align_backward_f32_range.argtypes = [fp, fp] + [ctypes.c_int] * 3 + [ctypes.c_float, ctypes.c_int, ctypes.c_int, fp] + \
        [ctypes.c_int] * 6
    lib.map_roi_levels_ref.argtypes = [fp, ctypes.c_int, ctypes.c_int, ctypes.c_float, ip]
    return lib


_L = None


def _get():
    global _L
    if _L is None:
        _L = _lib()
    return _L


def _f(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _i(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def roi_align_forward(feat, rois, output_size=7, spatial_scale=1.0, sampling_ratio=0, aligned=True, threads=1):
    """feat [N,C,H,W] float32, rois [R,5] -> (out [R,C,ph,pw], grid int32 [R,2], batch_idx int32 [R])."""
    feat = np.ascontiguousarray(feat, dtype=np.float32)
    rois = np.ascontiguousarray(rois, dtype=np.float32).reshape(-1, 5)
    N, C, H, W = feat.shape
    R = rois.shape[0]
    ph = pw = int(output_size)
    out = np.zeros((R, C, ph, pw), np.float32)
    grid = np.zeros((R, 2), np.int32)
    bidx = np.zeros((R,), np.int32)
    lib = _get()

    def run(r0, r1):
        lib.roi_align_forward_ref_range(_f(feat), N, C, H, W, _f(rois), r0, r1, ph, pw, float(spatial_scale),
                                        int(sampling_ratio), int(bool(aligned)), _f(out), _i(grid), _i(bidx))

    if threads <= 1 or R < 2 * threads:
        run(0, R)
    else:
        step = (R + threads - 1) // threads
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(lambda k: run(k * step, min(R, (k + 1) * step)), range(threads)))
    return out, grid, bidx


def _channel_pool(fn, C, threads):
    """Run fn(c0, c1) over disjoint channel ranges on a host thread pool (ctypes drops the GIL; every thread owns
    whole channel planes of the gradient, so no atomics are needed)."""
    if threads <= 1 or C < 2 * threads:
        fn(0, C)
        return
    step = (C + threads - 1) // threads
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(lambda k: fn(k * step, min(C, (k + 1) * step)), range((C + step - 1) // step)))


def roi_align_backward(gout, rois, feat_shape, output_size=7, spatial_scale=1.0, sampling_ratio=0, aligned=True, threads=None,
                       dtype=np.float64):
    """gout [R,C,ph,pw] -> grad_input [N,C,H,W].  dtype float64 (default): exact-order-free reference for the parity
    tests; float32: the reference's own arithmetic (fp32 `+=`), what bench.py's CPU baseline times."""
    gout = np.ascontiguousarray(gout, dtype=np.float32)
    rois = np.ascontiguousarray(rois, dtype=np.float32).reshape(-1, 5)
    N, C, H, W = feat_shape
    threads = threads or min(os.cpu_count() or 1, 64)
    gin = np.zeros((N, C, H, W), dtype)
    lib = _get()
    args = (_f(gout), _f(rois), rois.shape[0], int(output_size), int(output_size), float(spatial_scale), int(sampling_ratio),
            int(bool(aligned)))
    if dtype == np.float64:
        gp = gin.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        _channel_pool(lambda c0, c1: lib.roi_align_backward_ref_range(*args, gp, N, C, H, W, c0, c1), C, threads)
    else:
        gp = _f(gin)
        _channel_pool(lambda c0, c1: lib.roi_align_backward_f32_range(*args, gp, N, C, H, W, c0, c1), C, threads)
    return gin


def map_roi_levels(rois, num_levels, finest_scale=56.0):
    rois = np.ascontiguousarray(rois, dtype=np.float32).reshape(-1, 5)
    out = np.zeros((rois.shape[0],), np.int32)
    _get().map_roi_levels_ref(_f(rois), rois.shape[0], int(num_levels), float(finest_scale), _i(out))
    return out
