#!/usr/bin/env python3
"""Isolated attention timing with experiment switches (results are WRONG when a switch is set)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from foundpose_b200 import _native  # noqa: E402

B, N, H = 64, 901, 16
g = torch.Generator().manual_seed(0)
qkv = (torch.randn(B * N, 3 * H * 64, generator=g)).half().cuda()
lib = _native.load()
for rep in range(2):
    for name, fl in [("baseline", 0)]:
        lib.fp_gemm_force_1sm(ctypes.c_int(fl << 16))
        for _ in range(3):
            _native.attention_f16(qkv, B, N, H)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            _native.attention_f16(qkv, B, N, H)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 10 * 1e3
        print(f"rep{rep} {name:18s} {us:7.1f} us  {4.0 * B * H * N * N * 64 / us / 1e6:6.1f} TFLOP/s")
lib.fp_gemm_force_1sm(ctypes.c_int(0))
