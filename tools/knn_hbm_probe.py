#!/usr/bin/env python3
"""The k-NN kernel in its HBM-bound pass for ncu: 128 queries sweep the configs[2] bank (10.24 M x 384 fp16) once per
search, the bank split over all SMs + merge.  6 searches (ncu skips the warm-up launches with -s)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    from foundpose_b200 import _native
    from foundpose_b200.utils import knn_util

    dev = torch.device("cuda")
    rows, dim = int(os.environ.get("KNN_PROBE_TEMPLATES", "10000")) * 1024, 384
    bank = torch.empty(rows, dim, device=dev, dtype=torch.float16)
    for s0 in range(0, rows, 1 << 20):
        bank[s0:s0 + (1 << 20)] = torch.randn(min(1 << 20, rows - s0), dim, device=dev, dtype=torch.float16)
    knn = knn_util.KNN.from_packed(bank, _native.row_sqnorm_f16(bank), k=int(os.environ.get("KNN_PROBE_K", "5")), metric="l2")
    q = torch.randn(128, dim, device=dev)
    for _ in range(6):
        knn.search(q)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
