#!/usr/bin/env python3
"""A/B of the two residual epilogues of the pair GEMM (x += gamma * (A B^T + bias)):
mode 0 = TMA reduce-add (L2 atomics), mode 1 = TMA load of the x block + add in shared memory + TMA store.
Checks both against torch and times the proj / fc2 shapes of ViT-L at B=64."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from foundpose_b200 import _native  # noqa: E402

lib = _native.load()
M = 57664
g = torch.Generator().manual_seed(0)
for name, n, k in (("proj", 1024, 1024), ("fc2 ", 1024, 4096)):
    a = (torch.randn(M, k, generator=g)).half().cuda()
    b = (torch.randn(n, k, generator=g) * 0.03).half().cuda()
    bias = torch.randn(n, generator=g).cuda()
    gamma = torch.randn(n, generator=g).cuda()
    x0 = torch.randn(M, n, generator=g).cuda()
    ref = x0[:4096] + gamma * (a[:4096].float() @ b.float().t() + bias)
    for mode in (0, 1, 0, 1):
        lib.fp_gemm_force_1sm(ctypes.c_int(mode << 1))
        x = x0.clone()
        _native.gemm_tn_f16(a, b, _native.EPI_RESID_F32, bias=bias, gamma=gamma, out_f32=x)
        err = ((x[:4096] - ref).norm() / ref.norm()).item()
        tail_ok = torch.allclose(x[-300:], x0[-300:] + gamma * (a[-300:].float() @ b.float().t() + bias), rtol=2e-3, atol=2e-3)
        for _ in range(3):
            _native.gemm_tn_f16(a, b, _native.EPI_RESID_F32, bias=bias, gamma=gamma, out_f32=x)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            _native.gemm_tn_f16(a, b, _native.EPI_RESID_F32, bias=bias, gamma=gamma, out_f32=x)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 10 * 1e3
        print(f"{name} mode {mode}: {us:7.1f} us  {2.0 * M * n * k / us / 1e6:7.1f} TFLOP/s  rel_err {err:.2e}  tail_ok {tail_ok}", flush=True)
lib.fp_gemm_force_1sm(ctypes.c_int(0))
