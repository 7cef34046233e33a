"""Micro-batch size must not change a single output bit (GPU).

`bench.py` runs the 512 crops of a configs[2] step as micro-batches of 256 (fewer partial waves in the N = 1024 GEMMs
than at 64); the parity tests and the oracle spot check run smaller batches.  Every kernel on the path reduces in an
order that does not depend on how many crops share a launch (K loop of the GEMMs, one (image, head) per attention
item, row-wise LayerNorm statistics, per-crop k-NN items), so the outputs for crop c must be bit-identical whether it
travels in a batch of 256 or of 64 - the same property `tools/rank_equality_check.py` asserts across ranks.
Reference: the per-crop loop of scripts/infer.py:368-545 has no cross-crop state.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

FIELDS = ("template_ids", "count", "query_ids", "vertex_ids", "template_scores", "dists", "coord_2d", "coord_3d")


def _run(pipe, images, masks, B, dev):
    outs = {f: [] for f in FIELDS}
    for s in range(0, images.shape[0], B):
        out = pipe.run(images[s:s + B].to(dev), masks[s:s + B].to(dev))
        k = out.query_ids.shape[2]
        valid = torch.arange(k, device=dev).view(1, 1, k) < out.count.unsqueeze(-1)
        for f in FIELDS:
            t = getattr(out, f).clone()
            if t.dim() >= 3 and t.shape[2] == k:      # rows beyond count are scratch, not results
                t = torch.where(valid if t.dim() == 3 else valid.unsqueeze(-1), t, torch.zeros_like(t))
            outs[f].append(t)
    return {f: torch.cat(v) for f, v in outs.items()}


def test_vitl_pipeline_outputs_do_not_depend_on_the_micro_batch_size():
    import bench
    from foundpose_b200 import pipeline, synthetic
    from foundpose_b200.utils import dinov2_utils

    dev = torch.device("cuda", 0)
    wl = dict(bench.WORKLOADS["config3"], templates=300, patches=512)      # ViT-L/14 layer 9, PCA-384, 300-template bank
    arch, opts = bench.vit_arch_and_layer(wl["vit"])
    sd = synthetic.make_vit_state_dict(arch, seed=0, depth=opts["layer"] + 1)
    repre, _, _, _ = bench.build_repre_on_device(wl, dev, 0, 1)
    index = pipeline.ObjectIndex(repre, dev)
    n_total = 256
    images = synthetic.make_crops(n_total, (420, 420), seed=11)
    masks = synthetic.make_masks(n_total, (420, 420), seed=12).to(torch.uint8)
    results = {}
    for B in (256, 64):
        extractor = dinov2_utils.DinoFeatureExtractor(wl["vit"], state_dict=sd, max_batch=B).to(dev)
        pipe = pipeline.CropBatchPipeline(extractor, index, repre.feat_raw_projectors, B, crop_size=(420, 420),
                                          grid_cell_size=14.0, top_n_templates=5, top_k_buddies=300)
        results[B] = _run(pipe, images, masks, B, dev)
        del pipe, extractor
        torch.cuda.empty_cache()
    assert int(results[256]["count"].sum()) > 0
    for f in FIELDS:
        a, b = results[256][f], results[64][f]
        if a.is_floating_point():
            a, b = torch.nan_to_num(a), torch.nan_to_num(b)
        assert torch.equal(a, b), f"{f} differs between micro-batches of 256 and 64"
