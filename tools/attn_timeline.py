#!/usr/bin/env python3
"""Prints the event timeline of CTA 0 of the attention kernel (first items) from the debug buffer."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from foundpose_b200 import _native  # noqa: E402

B, N, H = 64, 901, 16
qkv = torch.randn(B * N, 3 * H * 64, generator=torch.Generator().manual_seed(0)).half().cuda()
lib = _native.load()
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
lib.fp_gemm_force_1sm(ctypes.c_int(flags << 16))
for _ in range(2):
    _native.attention_f16(qkv, B, N, H)
buf = torch.zeros(5 * 2048, dtype=torch.int64, device="cuda")
lib.fp_attention_debug_buffer(ctypes.c_void_p(buf.data_ptr()))
_native.attention_f16(qkv, B, N, H)
torch.cuda.synchronize()
lib.fp_attention_debug_buffer(ctypes.c_void_p(0))
ev = buf.cpu().view(5, 2048)
roles = ["producer", "S-issuer", "PV-issuer", "softmaxA", "softmaxB"]
rows = []
for r in range(5):
    for v in ev[r].tolist():
        if v == 0:
            continue
        rows.append((v & 0xffffffffffff, r, (v >> 48) & 0xffff))
rows.sort()
t0 = rows[0][0]
limit = int(sys.argv[1]) if len(sys.argv) > 1 else 140
for t, r, tag in rows[:limit]:
    print(f"{t - t0:8d} cyc {roles[r]:10s} tag={tag}")
# per-item duration seen by softmax A (tag 90 = drain)
drains = [t for t, r, tag in rows if r == 3 and tag == 90]
if len(drains) > 2:
    d = [b - a for a, b in zip(drains, drains[1:])]
    print("item period (SM cycles) seen by softmax A:", d[:12], "mean", sum(d) / len(d))
