#!/usr/bin/env python3
"""The four ViT-L GEMM shapes of a 64-crop micro-batch: the fused-epilogue tcgen05 kernels next to cuBLAS
(`torch.matmul` / `torch.addmm`, fp16 in, fp32 accumulate, NO fused GELU / residual / LayerNorm work) in the same
process, interleaved, after a warm-up long enough for the power cap to settle.  Context for `roofline.gemm` in the
bench line: how far the library GEMM itself gets on these shapes under the same conditions.

    python tools/gemm_vs_cublas.py [--iters 20] [--reps 3]          # prints one JSON line per (shape, implementation)
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from foundpose_b200 import _native  # noqa: E402

M = 64 * 901


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    g = torch.Generator().manual_seed(0)

    def rnd(*shape, scale=1.0):
        return (torch.randn(*shape, generator=g) * scale).half().cuda()

    shapes = [("qkv", 3072, 1024, _native.EPI_BIAS_F16), ("proj", 1024, 1024, _native.EPI_RESID_F32),
              ("fc1", 4096, 1024, _native.EPI_BIAS_GELU_F16), ("fc2", 1024, 4096, _native.EPI_RESID_F32)]
    bufs = {}
    for name, n, k, _ in shapes:
        bufs[name] = dict(a=rnd(M, k), b=rnd(n, k, scale=0.03), bias=torch.randn(n, generator=g).cuda(),
                          bias16=torch.randn(n, generator=g).half().cuda(), gamma=torch.randn(n, generator=g).cuda(),
                          o16=torch.empty((M, n), dtype=torch.float16, device="cuda"), o32=torch.zeros((M, n), device="cuda"))

    def ours(name, epi):
        b = bufs[name]
        if epi == _native.EPI_RESID_F32:
            _native.gemm_tn_f16(b["a"], b["b"], epi, bias=b["bias"], gamma=b["gamma"], out_f32=b["o32"])
        else:
            _native.gemm_tn_f16(b["a"], b["b"], epi, bias=b["bias"], out_f16=b["o16"])

    def cublas_plain(name, epi):
        b = bufs[name]
        torch.matmul(b["a"], b["b"].t(), out=b["o16"])

    def cublas_bias(name, epi):
        b = bufs[name]
        torch.addmm(b["bias16"], b["a"], b["b"].t(), out=b["o16"])

    impls = [("fused tcgen05 kernel (bias + GELU / LayerScale-residual epilogue)", ours),
             ("cuBLAS matmul, no epilogue", cublas_plain), ("cuBLAS addmm (bias only)", cublas_bias)]
    # warm-up: ~2 s of back-to-back GEMMs so that the clocks are the power-capped ones
    for _ in range(8):
        for name, n, k, epi in shapes:
            for _, fn in impls:
                for _ in range(5):
                    fn(name, epi)
    torch.cuda.synchronize()
    acc = {}
    for rep in range(args.reps):
        for name, n, k, epi in shapes:
            for label, fn in impls:
                fn(name, epi)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.iters):
                    fn(name, epi)
                e1.record()
                torch.cuda.synchronize()
                acc.setdefault((name, n, k, label), []).append(e0.elapsed_time(e1) / args.iters)
    for (name, n, k, label), ms in acc.items():
        best = sorted(ms)[len(ms) // 2]
        print(json.dumps({"gemm": name, "M": M, "N": n, "K": k, "impl": label, "us_median": round(best * 1e3, 1),
                          "tflops": round(2.0 * M * n * k / best / 1e9, 1), "us_all": [round(x * 1e3, 1) for x in ms]}))


if __name__ == "__main__":
    main()
