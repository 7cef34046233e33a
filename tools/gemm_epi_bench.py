#!/usr/bin/env python3
"""Isolated timing of the ViT-L GEMM shapes with each fused epilogue (same process, interleaved)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from foundpose_b200 import _native  # noqa: E402

M = 57664
g = torch.Generator().manual_seed(0)


def rnd(*shape, scale=1.0):
    return (torch.randn(*shape, generator=g) * scale).half().cuda()


import ctypes
cases = [ ("qkv  bias_f16", 3072, 1024, _native.EPI_BIAS_F16), ("qkv  nobias  ", 3072, 1024, -1),
         ("proj resid   ", 1024, 1024, _native.EPI_RESID_F32), ("proj bias_f16", 1024, 1024, _native.EPI_BIAS_F16),
         ("fc1  gelu    ", 4096, 1024, _native.EPI_BIAS_GELU_F16), ("fc1  bias_f16", 4096, 1024, _native.EPI_BIAS_F16),
         ("fc2  resid   ", 1024, 4096, _native.EPI_RESID_F32), ("fc2  bias_f16", 1024, 4096, _native.EPI_BIAS_F16)]
bufs = {}
for name, n, k, epi in cases:
    key = (n, k)
    if key not in bufs:
        bufs[key] = dict(a=rnd(M, k), b=rnd(n, k, scale=0.03), bias=torch.randn(n, generator=g).cuda(),
                         gamma=torch.randn(n, generator=g).cuda(), o16=torch.empty((M, n), dtype=torch.float16, device="cuda"),
                         o32=torch.zeros((M, n), device="cuda"))


def run(n, k, epi):
    b = bufs[(n, k)]
    if epi == 101:
        _native.load().fp_gemm_force_1sm(ctypes.c_int(4))
        _native.gemm_tn_f16(b["a"], b["b"], _native.EPI_BIAS_GELU_F16, bias=b["bias"], out_f16=b["o16"])
        _native.load().fp_gemm_force_1sm(ctypes.c_int(0))
    elif epi == 102:
        _native.load().fp_gemm_force_1sm(ctypes.c_int(8))
        _native.gemm_tn_f16(b["a"], b["b"], _native.EPI_BIAS_F16, bias=b["bias"], out_f16=b["o16"])
        _native.load().fp_gemm_force_1sm(ctypes.c_int(0))
    elif epi == -1:
        _native.gemm_tn_f16(b["a"], b["b"], _native.EPI_BIAS_F16, out_f16=b["o16"])
    elif epi == _native.EPI_RESID_F32:
        _native.gemm_tn_f16(b["a"], b["b"], epi, bias=b["bias"], gamma=b["gamma"], out_f32=b["o32"])
    else:
        _native.gemm_tn_f16(b["a"], b["b"], epi, bias=b["bias"], out_f16=b["o16"])


for rep in range(2):
    for name, n, k, epi in cases:
        for _ in range(3):
            run(n, k, epi)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            run(n, k, epi)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"rep{rep} {name} N={n:4d} K={k:4d}: {ms * 1e3:7.1f} us  {2.0 * M * n * k / ms / 1e9:7.1f} TFLOP/s")
