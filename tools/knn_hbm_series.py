#!/usr/bin/env python3
"""Time series of the HBM-bound k-NN pass (128 queries sweep a T x 1024 x d bank, k = 5): device time of every search of
a back-to-back run, with nvidia-smi power / clocks sampled alongside - is the pass power-capped when it runs for long?

    python tools/knn_hbm_series.py [--templates 10000] [--dim 384] [--iters 300]
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--templates", type=int, default=10000)
    ap.add_argument("--dim", type=int, default=384)
    ap.add_argument("--iters", type=int, default=300)
    ap.add_argument("--k", type=int, default=5)
    args = ap.parse_args()
    from foundpose_b200 import _native
    from foundpose_b200.utils import knn_util

    lib = _native.load()
    dev = torch.device("cuda")
    rows = args.templates * 1024
    bank = torch.empty(rows, args.dim, device=dev, dtype=torch.float16)
    for s0 in range(0, rows, 1 << 20):
        bank[s0:s0 + (1 << 20)] = torch.randn(min(1 << 20, rows - s0), args.dim, device=dev, dtype=torch.float16)
    knn = knn_util.KNN.from_packed(bank, _native.row_sqnorm_f16(bank), k=args.k, metric="l2")
    q = torch.randn(128, args.dim, device=dev)
    knn.search(q)
    torch.cuda.synchronize()
    time.sleep(3.0)   # let the GPU go idle, as after a CPU phase

    samples, stop = [], threading.Event()

    def sample():
        while not stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw", "--format=csv,noheader,nounits",
                                      "-i", "0"], capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                samples.append((time.perf_counter(), float(out[0]), float(out[1]), float(out[2])))
            except Exception:
                pass
            time.sleep(0.02)

    th = threading.Thread(target=sample, daemon=True)
    th.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.iters)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.iters)]
    t0 = time.perf_counter()
    for i in range(args.iters):
        starts[i].record()
        knn.search(q)
        ends[i].record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    stop.set()
    th.join()
    ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    gb = rows * args.dim * 2 / 1e9
    during = [s for s in samples if t0 <= s[0] <= t1]
    print(json.dumps({"templates": args.templates, "dim": args.dim, "k": args.k, "bank_gb": gb, "iters": args.iters,
                      "wall_s": round(t1 - t0, 3),
                      "search_ms_first10": [round(x, 3) for x in ms[:10]],
                      "search_ms_every_20th": [round(x, 3) for x in ms[::20]],
                      "gbs_best": round(gb / min(ms) * 1e3, 1), "gbs_last_quarter": round(gb / (sum(ms[-len(ms) // 4:]) / (len(ms) // 4)) * 1e3, 1),
                      "smi_during": [(round(s[0] - t0, 2), s[1], s[2], s[3]) for s in during][:40]}))


if __name__ == "__main__":
    main()
