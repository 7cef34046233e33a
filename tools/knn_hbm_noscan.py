import ctypes, sys
sys.path.insert(0, '/root/repo')
import torch
from foundpose_b200 import _native
from foundpose_b200.utils import knn_util
lib = _native.load(); dev = torch.device('cuda')
for T in (1000, 2000, 10000):
    rows, dim = T * 1024, 384
    bank = torch.randn(rows, dim, device=dev, dtype=torch.float16)
    bn = _native.row_sqnorm_f16(bank)
    q = torch.randn(128, dim, device=dev)
    for k in (5, 1):
        for fl in (0, 1):
            lib.fp_knn_set_flags(fl)
            index = knn_util.KNN.from_packed(bank, bn, k=k, metric='l2')
            for _ in range(3): index.search(q)
            torch.cuda.synchronize()
            for c in range(8): lib.fp_profile_read(ctypes.c_int(c), None, None, None, ctypes.c_int(1))
            lib.fp_profile_enable(1)
            for _ in range(10): index.search(q)
            torch.cuda.synchronize(); lib.fp_profile_enable(0)
            ms = ctypes.c_double(); lib.fp_profile_read(ctypes.c_int(4), ctypes.byref(ms), None, None, ctypes.c_int(1))
            t = ms.value / 10 * 1e-3
            print(T, 'k', k, 'noscan' if fl else 'scan', round(t * 1e6, 1), 'us', round(rows * dim * 2 / t / 1e9 / 6449.4, 3))
    lib.fp_knn_set_flags(0)
