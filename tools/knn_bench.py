#!/usr/bin/env python3
"""k-NN kernel benchmark: HBM-bound (few queries, split bank) and tensor-bound (K4, many queries) regimes.

Prints one JSON line per case: achieved GB/s of bank bytes (F*d*2 per sweep) and TFLOP/s (2*Nq*F*d),
with the fraction of the measured peaks in MEASURED_PEAKS.json.  BASELINE.json configs 3 and 5.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from foundpose_b200 import _native  # noqa: E402
from foundpose_b200.utils import knn_util  # noqa: E402

peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
HBM = float(peaks.get("hbm_gbs", 6650.0))
TC = float(peaks.get("bf16_tflops", 1590.0))     # burst: kernel timed alone
dev = torch.device("cuda", 0)


def make_index(rows, dim, k):
    bank = torch.empty(rows, dim, device=dev, dtype=torch.float16)
    step = 1 << 20
    for s0 in range(0, rows, step):        # bounded temporaries
        bank[s0:s0 + step] = torch.randn(min(step, rows - s0), dim, device=dev, dtype=torch.float16)
    return knn_util.KNN.from_packed(bank, _native.row_sqnorm_f16(bank), k=k, metric="l2")


def run(tag, nq, templates, patches, dim, k, iters=5):
    rows = templates * patches
    index = make_index(rows, dim, k)
    q = torch.randn(nq, dim, device=dev)
    for _ in range(2):
        index.search(q)
    torch.cuda.synchronize()
    lib = _native.load()
    import ctypes
    lib.fp_profile_enable(1)
    for _ in range(iters):
        index.search(q)
    lib.fp_profile_enable(0)
    ms, w, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_int()
    lib.fp_profile_read(ctypes.c_int(4), ctypes.byref(ms), ctypes.byref(w), ctypes.byref(n), ctypes.c_int(1))
    for c in (5, 6):
        lib.fp_profile_read(ctypes.c_int(c), None, None, None, ctypes.c_int(1))
    t = ms.value / iters * 1e-3            # knn_kernel device time per search
    gbs = rows * dim * 2 / t / 1e9
    tf = 2.0 * nq * rows * dim / t / 1e12
    print(json.dumps({"case": tag, "queries": nq, "templates": templates, "patches": patches, "dim": dim, "k": k,
                      "bank_gb": rows * dim * 2 / 1e9, "knn_kernel_ms": t * 1e3, "bank_gbs": round(gbs, 1),
                      "frac_hbm_peak": round(gbs / HBM, 3), "tflops": round(tf, 1),
                      "frac_tensor_peak": round(tf / TC, 3),
                      "bound": "hbm" if gbs / HBM > tf / TC else "tensor"}), flush=True)
    del index, q
    torch.cuda.empty_cache()


if __name__ == "__main__":
    # HBM-bound regime: one 128-query tile sweeps the whole bank once (bank split over all SMs).
    for t in (1000, 2000, 5000, 10000, 20000, 50000):
        for d in (384, 768):
            if t * 1024 * d * 2 > 45e9:   # keep the sweep well inside the 180 GB of HBM
                continue
            run("hbm_sweep", 128, t, 1024, d, 5)
    # Tensor-bound regime (K4, BASELINE configs 3/5): every crop's 900 queries vs the full bank.
    run("k4_config2_bank", 64 * 900, 2000, 1024, 256, 5, iters=2)
    run("k4_config3_8crops", 8 * 900, 10000, 1024, 384, 5, iters=2)
