#!/usr/bin/env python3
"""Variants of the full-bank k-NN (K4) on one micro-batch of queries, to see what bounds the pair kernel.

    python tools/k4_probe.py [--templates 10000] [--crops 64] [--dim 384]
Prints one line per variant: milliseconds and TFLOP/s (2 * nq * F * d flops).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--templates", type=int, default=10000)
    ap.add_argument("--crops", type=int, default=64)
    ap.add_argument("--dim", type=int, default=384)
    ap.add_argument("--variants", type=str, default="pair,pair_noscan,pair_stream,one_cta,pair_l2")
    ap.add_argument("--iters", type=int, default=4)
    args = ap.parse_args()
    from foundpose_b200 import _native
    from foundpose_b200.utils import knn_util

    lib = _native.load()
    dev = torch.device("cuda")
    F, d, nq = args.templates * 1024, args.dim, args.crops * 900
    g = torch.Generator(device=dev).manual_seed(0)
    bank = torch.empty((F, d), dtype=torch.float16, device=dev)
    for s in range(0, F, 1 << 20):
        n = min(1 << 20, F - s)
        bank[s:s + n] = torch.randn((n, d), generator=g, device=dev).to(torch.float16)
    bn = _native.row_sqnorm_f16(bank)
    q = torch.randn((nq, d), generator=g, device=dev).to(torch.float16)
    qn = _native.row_sqnorm_f16(q)
    index = knn_util.KNN.from_packed(bank, bn, k=5, metric="l2")
    small = knn_util.KNN.from_packed(bank[:65536], bn[:65536], k=5, metric="l2")   # 50 MB: stays in L2

    import threading

    try:
        import pynvml

        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(0)
    except Exception:
        pynvml = handle = None
    power = {}

    def time_it(fn, flops):
        fn()
        torch.cuda.synchronize()
        samples, stop = [], threading.Event()

        def sample():
            while not stop.is_set() and pynvml is not None:
                samples.append((pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM),
                                pynvml.nvmlDeviceGetPowerUsage(handle) / 1000.0))
                stop.wait(0.02)

        th = threading.Thread(target=sample, daemon=True)
        th.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        stop.set()
        th.join()
        ms = e0.elapsed_time(e1) / args.iters
        if samples:
            samples.sort()
            power["sm_mhz_median"] = samples[len(samples) // 2][0]
            power["watts_median"] = sorted(w for _, w in samples)[len(samples) // 2]
        return ms, flops / (ms * 1e-3) / 1e12

    def one_cta():
        n_items = _native.knn_num_items(nq)
        items = _native.new_knn_items(n_items, dev)
        _native.knn_items_dense(items, nq, 0, F)
        dd = torch.empty((nq, 5), dtype=torch.float32, device=dev)
        ii = torch.empty((nq, 5), dtype=torch.int64, device=dev)
        _native.knn_search_items(q, qn, bank, bn, items, n_items, 0, 5, dd, ii)

    full = 2.0 * nq * F * d
    for v in args.variants.split(","):
        lib.fp_knn_set_flags(0)
        if v == "pair":
            ms, tf = time_it(lambda: index.search_packed(q, qn), full)
        elif v == "pair_noscan":
            lib.fp_knn_set_flags(1)
            ms, tf = time_it(lambda: index.search_packed(q, qn), full)
        elif v == "pair_stream":
            lib.fp_knn_set_flags(2)
            ms, tf = time_it(lambda: index.search_packed(q, qn), full)
        elif v == "pair_stream_noscan":
            lib.fp_knn_set_flags(3)
            ms, tf = time_it(lambda: index.search_packed(q, qn), full)
        elif v == "pair_nosync":
            lib.fp_knn_set_flags(8)
            ms, tf = time_it(lambda: index.search_packed(q, qn), full)
        elif v == "pair_ldtm":   # TMEM reads only: needs a library built with `make EXTRA=-DFP_KNN_EXPERIMENTS`
            lib.fp_knn_set_flags(4)
            ms, tf = time_it(lambda: index.search_packed(q, qn), full)
        elif v == "one_cta":
            ms, tf = time_it(one_cta, full)
        elif v == "pair_l2":
            ms, tf = time_it(lambda: small.search_packed(q, qn), 2.0 * nq * 65536 * d)
        else:
            continue
        lib.fp_knn_set_flags(0)
        print(json.dumps({"variant": v, "ms": round(ms, 3), "tflops": round(tf, 1), "nq": nq, "bank_rows": F, "dim": d,
                          **power}))


if __name__ == "__main__":
    main()
