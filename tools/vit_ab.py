#!/usr/bin/env python3
"""A/B timing of the ViT stage under the GEMM tuning switches, same process / same GPU."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from foundpose_b200 import _native, synthetic  # noqa: E402
from foundpose_b200.utils import dinov2_utils  # noqa: E402

dev = torch.device("cuda", 0)
lib = _native.load()
arch = synthetic.VIT_ARCHS["vitl14"]
sd = synthetic.make_vit_state_dict(arch, seed=0, depth=10)
B = 64
ext = dinov2_utils.DinoFeatureExtractor("dinov2_vitl14", state_dict=sd, max_batch=B).to(dev)
imgs = synthetic.make_crops(B, (420, 420), seed=1).to(dev)
out = torch.empty((B, 900, 1024), device=dev)
names = ["gemm", "attention", "layernorm", "vit_misc"]


def run(flags, steps=8):
    lib.fp_gemm_force_1sm(ctypes.c_int(flags))
    for _ in range(3):
        ext.forward_tokens(imgs, want_cls=False, out_tokens=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        ext.forward_tokens(imgs, want_cls=False, out_tokens=out)
    e1.record()
    torch.cuda.synchronize()
    total = e0.elapsed_time(e1) / steps
    lib.fp_profile_enable(1)
    for _ in range(3):
        ext.forward_tokens(imgs, want_cls=False, out_tokens=out)
    lib.fp_profile_enable(0)
    fam = {}
    for c, n in enumerate(names):
        ms, w, k = ctypes.c_double(), ctypes.c_double(), ctypes.c_int()
        lib.fp_profile_read(ctypes.c_int(c), ctypes.byref(ms), ctypes.byref(w), ctypes.byref(k), ctypes.c_int(1))
        fam[n] = round(ms.value / 3, 3)
    return total, fam


configs = {"2sm+prefetch+fastgelu": 0, "1sm": 1, "2sm noprefetch": 2, "2sm libm gelu": 4, "2sm noprefetch libm": 6}
for rep in range(2):
    for name, flags in configs.items():
        t, fam = run(flags)
        print(f"rep{rep} {name:24s} vit={t:7.3f} ms  {fam}")
