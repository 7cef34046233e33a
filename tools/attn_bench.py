#!/usr/bin/env python3
"""Isolated attention timing over the launcher's kernel variants (share of exponentials on the FMA
pipe), each checked against an fp32 torch softmax(QK^T)V on the same fp16 inputs."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from foundpose_b200 import _native  # noqa: E402

B, N, H = 64, 901, 16
# flags: bits 0-7 = kernel variant + 1 (0 = production), bits 8-13 = start skew of query tile B in units of 128 cycles,
# bit 14 = round-1 issue split (S issuer / PV issuer) instead of one issuer warp per query tile
OLD = 1 << 14
VARIANTS = [("r1 split 24|24", 1 | OLD), ("per-tile skew0", 1), ("per-tile skew512", 1 | (4 << 8)),
            ("per-tile skew1024", 1 | (8 << 8)), ("per-tile skew1536", 1 | (12 << 8)), ("per-tile skew2048", 1 | (16 << 8)),
            ("per-tile skew3072", 1 | (24 << 8)), ("per-tile 32|32 s1024", 4 | (8 << 8)), ("per-tile 16|16 s1024", 3 | (8 << 8)),
            ("2thr/row 0|0 per-tile", 5), ("2thr/row 16|16 per-tile", 6), ("2thr/row 8|24 per-tile", 7),
            ("2thr/row 16|16 r1 split", 6 | OLD), ("default", 0)]
if len(sys.argv) > 1:
    VARIANTS = [v for v in VARIANTS if any(a in v[0] for a in sys.argv[1:])]
g = torch.Generator().manual_seed(0)
qkv = (torch.randn(B * N, 3 * H * 64, generator=g) * 1.5).half().cuda()
lib = _native.load()


def reference(b):
    x = qkv[b * N:(b + 1) * N].float().view(N, 3, H, 64).permute(1, 2, 0, 3)   # [3, H, N, 64]
    att = torch.softmax(x[0] @ x[1].transpose(1, 2) * 0.125, dim=-1)
    return (att @ x[2]).permute(1, 0, 2).reshape(N, H * 64)


ref0 = reference(0)
for rep in range(2):
    for name, fl in VARIANTS:
        lib.fp_gemm_force_1sm(ctypes.c_int(fl << 16))
        for _ in range(3):
            out = _native.attention_f16(qkv, B, N, H)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            _native.attention_f16(qkv, B, N, H)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 10 * 1e3
        err = ((out[:N].float() - ref0).norm() / ref0.norm()).item()
        print(f"rep{rep} {name:16s} {us:7.1f} us  {4.0 * B * H * N * N * 64 / us / 1e6:6.1f} TFLOP/s  rel_err {err:.2e}", flush=True)
lib.fp_gemm_force_1sm(ctypes.c_int(0))
