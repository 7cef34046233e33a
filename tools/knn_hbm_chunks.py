#!/usr/bin/env python3
"""HBM-bound k-NN pass on small banks: device time of knn_kernel<5> (128 queries, d = 384) against the number of bank
chunks the search is split into (round-2 finding: one chunk per SM = 148 is the optimum for 1 k / 2 k / 5 k templates;
every further item per SM costs a pipeline restart - profiles/r02_knn_hbm_chunks.md)."""
import ctypes, json, sys
sys.path.insert(0, '/root/repo')
import torch
from foundpose_b200 import _native
from foundpose_b200.utils import knn_util
lib = _native.load(); dev = torch.device('cuda')
for T in (1000, 2000, 5000):
    rows, dim = T * 1024, 384
    bank = torch.randn(rows, dim, device=dev, dtype=torch.float16)
    bn = _native.row_sqnorm_f16(bank)
    q = torch.randn(128, dim, device=dev)
    for want in (74, 148, 222, 296, 444, 592):
        knn_util._choose_num_chunks = lambda n_q, w=want: w
        index = knn_util.KNN.from_packed(bank, bn, k=5, metric='l2')
        for _ in range(3): index.search(q)
        torch.cuda.synchronize()
        for c in range(8): lib.fp_profile_read(ctypes.c_int(c), None, None, None, ctypes.c_int(1))
        lib.fp_profile_enable(1)
        for _ in range(10): index.search(q)
        torch.cuda.synchronize(); lib.fp_profile_enable(0)
        ms = ctypes.c_double(); lib.fp_profile_read(ctypes.c_int(4), ctypes.byref(ms), None, None, ctypes.c_int(1))
        t = ms.value / 10 * 1e-3
        print(T, want, round(t * 1e6, 1), 'us', round(rows * dim * 2 / t / 1e9), 'GB/s', round(rows * dim * 2 / t / 1e9 / 6449.4, 3))
