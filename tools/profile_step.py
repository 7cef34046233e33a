#!/usr/bin/env python3
"""Runs micro-batches of a bench workload inside a cudaProfilerStart/Stop range (for ncu).

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py --micro 1
    ncu --profile-from-start off --set full --clock-control none --import-source on \
        -o gpurun_out/prof_step python tools/profile_step.py --micro 1

One micro-batch = `batch` crops of the workload (config3: 256) through the whole path, incl. the full-bank search K4
(launched per 64 crops).
Numbers printed under ncu are never bench values; this script prints none.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="config3")
    ap.add_argument("--micro", type=int, default=1, help="micro-batches inside the profiled range")
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--no-k4", action="store_true")
    args = ap.parse_args()
    from foundpose_b200 import pipeline, synthetic
    from foundpose_b200.utils import dinov2_utils, knn_util

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    wl = bench.WORKLOADS[args.workload]
    B = wl["batch"]
    arch, opts = bench.vit_arch_and_layer(wl["vit"])
    sd = synthetic.make_vit_state_dict(arch, seed=0, depth=opts["layer"] + 1)
    extractor = dinov2_utils.DinoFeatureExtractor(wl["vit"], state_dict=sd, max_batch=B).to(dev)
    repre, bank, pdict, _ = bench.build_repre_on_device(wl, dev, 0, 1)
    index = pipeline.ObjectIndex(repre, dev)
    pipe = pipeline.CropBatchPipeline(extractor, index, repre.feat_raw_projectors, B, top_n_templates=wl["top_n"],
                                      top_k_buddies=wl["top_k"])
    k4 = 0 if args.no_k4 else wl["k4"]
    k4_index = knn_util.KNN.from_packed(index.bank16, index.bank_sqnorm, k=k4, metric="l2") if k4 else None
    images = [synthetic.make_crops(B, (420, 420), seed=100 + s).to(dev) for s in range(2)]
    masks = torch.ones(B, 420, 420, dtype=torch.uint8, device=dev)

    k4_rows = min(B, 64) * pipe.stride     # the full-bank search is launched per 64 crops, as in bench.py

    def micro(i: int) -> None:
        pipe.run(images[i % 2], masks)
        if k4_index is not None:
            for r0 in range(0, B * pipe.stride, k4_rows):
                k4_index.search_packed(pipe.proj16[r0:r0 + k4_rows], pipe.engine.q_sqnorm[r0:r0 + k4_rows])

    for i in range(args.warmup):
        micro(i)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for i in range(args.micro):
        micro(i)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
