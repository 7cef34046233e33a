#!/usr/bin/env python3
"""Runs a few steps of the bench workload inside a cudaProfilerStart/Stop range (for ncu).

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py --steps 1
    ncu --profile-from-start off --set full --clock-control none --import-source on \
        -k regex:gemm_tn_kernel -c 6 -o gpurun_out/prof_gemm python tools/profile_step.py --steps 1

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
    ap.add_argument("--workload", default="config2")
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    from foundpose_b200 import pipeline, synthetic
    from foundpose_b200.utils import dinov2_utils, knn_util, projector_util, repre_util, template_util

    dev = torch.device("cuda", 0)
    wl = bench.WORKLOADS[args.workload]
    B = wl["batch"]
    arch, opts = bench.vit_arch_and_layer(wl["vit"])
    sd = synthetic.make_vit_state_dict(arch, seed=0, depth=opts["layer"] + 1)
    extractor = dinov2_utils.DinoFeatureExtractor(wl["vit"], state_dict=sd, max_batch=B).to(dev)
    pdict = synthetic.make_pca(arch.embed_dim, wl["dim"], seed=0)
    projectors = [projector_util.projector_from_tensordict(pdict)]
    bank = bench.build_bank_cpu(wl)
    feat, centroids = bank["feat_vectors"].to(dev), bank["feat_cluster_centroids"].to(dev)
    tpl_ids = bank["feat_to_template_ids"].to(dev)
    wk = knn_util.KNN(k=1, metric="l2")
    wk.fit(centroids)
    f2w = wk.search(feat)[1].flatten()
    descs, idfs = template_util.calc_tfidf_descriptors(feat, f2w, tpl_ids, centroids, wl["templates"], 3, False, 10.0)
    repre = repre_util.FeatureBasedObjectRepre(
        vertices=bank["vertices"].to(dev), feat_vectors=feat, feat_to_template_ids=tpl_ids,
        feat_cluster_centroids=centroids, feat_cluster_idfs=idfs, template_descs=descs,
        template_desc_opts=repre_util.TemplateDescOpts(), feat_raw_projectors=projectors)
    index = pipeline.ObjectIndex(repre, dev)
    pipe = pipeline.CropBatchPipeline(extractor, index, projectors, B, top_n_templates=wl["top_n"],
                                      top_k_buddies=wl["top_k"])
    images = [synthetic.make_crops(B, (420, 420), seed=100 + s).to(dev) for s in range(2)]
    masks = torch.ones(B, 420, 420, dtype=torch.uint8, device=dev)
    for i in range(args.warmup):
        pipe.run(images[i % 2], masks)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for i in range(args.steps):
        pipe.run(images[i % 2], masks)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
