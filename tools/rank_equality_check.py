#!/usr/bin/env python3
"""Multi-GPU result equality: the outputs rank r computes for its shard of the crops == the outputs a single GPU
computes for the same crops.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tools/rank_equality_check.py

Rank 0 builds a synthetic object representation and replicates it with distributed.broadcast_object_repre (NCCL);
every rank runs its contiguous shard of 2*B*world crops through CropBatchPipeline; rank 0 additionally runs ALL
crops alone.  Integer outputs (template ids, counts, 2D / 3D ids) must be bit-identical, float outputs equal.
Prints one JSON line on rank 0; exit code 1 on any mismatch.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main() -> int:
    import bench
    from foundpose_b200 import distributed, pipeline, synthetic
    from foundpose_b200.utils import dinov2_utils

    rank, world, local_rank = distributed.init_from_env()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    wl = dict(bench.WORKLOADS["config2"], templates=200, patches=512, vit="dinov2_version=vits14-reg_stride=14_"
              "facet=token_layer=9_norm=1")
    B = 8
    arch, opts = bench.vit_arch_and_layer(wl["vit"])
    sd = synthetic.make_vit_state_dict(arch, seed=0, depth=opts["layer"] + 1)
    extractor = dinov2_utils.DinoFeatureExtractor(wl["vit"], state_dict=sd, max_batch=B).to(dev)
    repre, _, _, t_bcast = bench.build_repre_on_device(wl, dev, rank, world)
    index = pipeline.ObjectIndex(repre, dev)
    pipe = pipeline.CropBatchPipeline(extractor, index, repre.feat_raw_projectors, B, crop_size=(420, 420),
                                      grid_cell_size=14.0, top_n_templates=5, top_k_buddies=300)
    n_total = 2 * B * world
    images = synthetic.make_crops(n_total, (420, 420), seed=7)
    masks = synthetic.make_masks(n_total, (420, 420), seed=8).to(torch.uint8)
    fields = ("template_ids", "count", "query_ids", "vertex_ids", "template_scores", "dists", "coord_2d", "coord_3d")

    def run(lo: int, hi: int):
        outs = {f: [] for f in fields}
        for s in range(lo, hi, B):
            out = pipe.run(images[s:s + B].to(dev), masks[s:s + B].to(dev))
            # rows at and beyond count[b, j] of the [B, topn, K, ...] buffers are not part of the result (they keep
            # whatever an earlier batch left there): blank them so that only results are compared
            k = out.query_ids.shape[2]
            valid = torch.arange(k, device=dev).view(1, 1, k) < out.count.unsqueeze(-1)
            for f in fields:
                t = getattr(out, f).clone()
                if t.dim() >= 3 and t.shape[2] == k:
                    t = torch.where(valid if t.dim() == 3 else valid.unsqueeze(-1), t, torch.zeros_like(t))
                outs[f].append(t)
        return {f: torch.cat(v) for f, v in outs.items()}

    lo, hi = distributed.shard_range(n_total, rank, world)
    mine = run(lo, hi)
    ok, report = True, {}
    for f in fields:
        gathered = [torch.empty_like(mine[f]) for _ in range(world)] if rank == 0 else None
        if world > 1:
            dist.gather(mine[f], gathered, dst=0)
        else:
            gathered = [mine[f]]
        if rank == 0:
            report[f] = torch.cat(gathered)
    if rank == 0:
        single = run(0, n_total)
        for f in fields:
            a, b = report[f], single[f]
            same = torch.equal(a, b) if not a.is_floating_point() else torch.equal(torch.nan_to_num(a), torch.nan_to_num(b))
            report[f] = bool(same)
            ok &= bool(same)
        print(json.dumps({"check": "rank outputs == single-GPU outputs", "world": world, "crops": n_total,
                          "bank_broadcast_s": t_bcast, "fields_equal": report, "ok": ok}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
