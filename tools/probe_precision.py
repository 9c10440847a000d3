#!/usr/bin/env python
"""Error of the dense engines against fp64 as a function of the reduction length K (GPU probe, not a test).

What it answers: does the fp32 accumulator of tcgen05.mma (TMEM) round to nearest or truncate, i.e. does the error of
the split-precision engines grow ~sqrt(K) (rounding noise) or ~K (bias)?  Two data sets per K: zero-mean products
(random signs) and all-positive products (worst case for a truncating accumulator).
    python tools/probe_precision.py > gpurun_out/precision.txt
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import unsupervised_domain_adaptation_object_detection_implementation_b200 as uda  # noqa: E402
from unsupervised_domain_adaptation_object_detection_implementation_b200 import functional as F_  # noqa: E402

dev = torch.device("cuda:0")
M, N = 256, 256
g = torch.Generator(device=dev).manual_seed(0)
print(f"{'K':>7} {'data':>9} {'engine':>12} {'max|err|/max|y|':>16} {'mean(err)/mean|y|':>18} {'rms(err)/rms(y)':>16}")
for K in (1024, 9216, 100352):
    for data in ("zero-mean", "positive"):
        x = torch.randn(M, K, device=dev, generator=g)
        w = torch.randn(N, K, device=dev, generator=g) * K ** -0.5
        if data == "positive":
            x, w = x.abs(), w.abs()
        ref = x.double() @ w.double().t()
        for engine in ("simt_f32", "umma_bf16x6", "umma_bf16x3", "umma_bf16"):
            xe = x.to(F_.act_dtype(engine)).view(M, 1, 1, K)
            we = w if engine != "umma_bf16" else w.to(torch.bfloat16)
            y = F_.dense_layer(xe, we, engine=engine, out_dtype=torch.float32).view(M, N).double()
            err = y - ref
            print(f"{K:7d} {data:>9} {engine:>12} {float(err.abs().max() / ref.abs().max()):16.3e} "
                  f"{float(err.mean() / ref.abs().mean()):18.3e} {float(err.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()):16.3e}")
        # torch's own fp32 GEMM (cuBLAS, TF32 off) for scale
        y = (x @ w.t()).double()
        err = y - ref
        print(f"{K:7d} {data:>9} {'cublas_f32':>12} {float(err.abs().max() / ref.abs().max()):16.3e} "
              f"{float(err.mean() / ref.abs().mean()):18.3e} {float(err.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()):16.3e}")
