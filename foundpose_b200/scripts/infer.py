#!/usr/bin/env python3
"""Entry point mirroring the reference's scripts/infer.py for the per-crop hot path.

`InferOpts` has the same fields and defaults as the reference (scripts/infer.py:55-100) and is read
from `--opts-path <json>` under the key "infer_opts" (utils/config_util.py:254-278) or from flags.
The reference's dataset plumbing (BOP images, CNOS detections, evaluation, rendering) is outside the
scope of this build; what this script runs is the per-instance block
scripts/infer.py:467-604 - extractor -> mask filter -> sampling -> PCA -> establish_correspondences ->
coarse pose per retrieved template (PnP-RANSAC) -> best coarse pose - on crops you provide (`--crops <file.pt>` with tensors "images" Bx3xHxW in [0,1] and "masks" BxHxW)
against a `repre.pth` object representation (`--repre-dir`), or on seeded synthetic crops and a
synthetic representation when neither is given (`--synthetic`).

Two execution modes produce the same outputs:
  * per crop, through the reference-compatible functions (feature_util / projector_util /
    corresp_util), exactly as the reference's loop body reads;
  * batched (`--batch N`), through pipeline.CropBatchPipeline (no host synchronisation per crop).
"""

from __future__ import annotations

import argparse
import json
import os
import sys
from typing import Any, Dict, List, NamedTuple, Optional, Tuple

import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from foundpose_b200 import distributed, pipeline, synthetic  # noqa: E402
from foundpose_b200.utils import (corresp_util, feature_util, knn_util, logging, misc,  # noqa: E402
                                  pnp_util, projector_util, repre_util, structs, template_util)


class InferOpts(NamedTuple):
    """Options that can be specified via the command line (reference scripts/infer.py:55-100)."""

    version: str = "v1"
    repre_version: str = "v1"
    object_dataset: str = "lmo"
    object_lids: Optional[List[int]] = None
    max_sym_disc_step: float = 0.01

    # Cropping options.
    crop: bool = True
    crop_rel_pad: float = 0.2
    crop_size: Tuple[int, int] = (420, 420)

    # Object instance options.
    use_detections: bool = True
    num_preds_factor: float = 1.0
    min_visibility: float = 0.1

    # Feature extraction options.
    extractor_name: str = "dinov2_vitl14"
    grid_cell_size: float = 1.0
    max_num_queries: int = 1000000

    # Feature matching options.
    match_template_type: str = "tfidf"
    match_top_n_templates: int = 5
    match_feat_matching_type: str = "cyclic_buddies"
    match_top_k_buddies: int = 300

    # PnP options.
    pnp_type: str = "opencv"
    pnp_ransac_iter: int = 1000
    pnp_required_ransac_conf: float = 0.99
    pnp_inlier_thresh: float = 10.0
    pnp_refine_lm: bool = True

    final_pose_type: str = "best_coarse"

    # Other options.
    save_estimates: bool = True
    vis_results: bool = True
    vis_corresp_top_n: int = 100
    vis_feat_map: bool = True
    vis_for_paper: bool = True
    debug: bool = True


def load_opts(path: Optional[str], overrides: Dict[str, Any]) -> InferOpts:
    values: Dict[str, Any] = {}
    if path:
        with open(path, "r") as f:
            values = dict(json.load(f)["infer_opts"])
    values.update({k: v for k, v in overrides.items() if v is not None})
    unknown = set(values) - set(InferOpts._fields)
    if unknown:
        raise ValueError(f"Unknown InferOpts fields: {sorted(unknown)}")
    if "crop_size" in values:
        values["crop_size"] = tuple(values["crop_size"])
    return InferOpts(**values)


def build_indices(repre: repre_util.FeatureBasedObjectRepre, opts: InferOpts, device: torch.device):
    """Index construction of scripts/infer.py:215-239 on the packed ObjectIndex (zero-copy views)."""
    index = pipeline.get_object_index(repre, device)
    visual_words_knn_index = None
    if opts.match_template_type == "tfidf":
        visual_words_knn_index = knn_util.KNN.from_packed(
            index.centroids16, index.centroid_sqnorm, k=repre.template_desc_opts.tfidf_knn_k,
            metric=repre.template_desc_opts.tfidf_knn_metric)
    template_knn_indices = []
    if opts.match_feat_matching_type == "cyclic_buddies":
        off = index.tpl_off.tolist()
        for t in range(index.num_templates):
            template_knn_indices.append(knn_util.KNN.from_packed(
                index.bank16[off[t]:off[t + 1]], index.bank_sqnorm[off[t]:off[t + 1]], k=1, metric="l2"))
    return index, visual_words_knn_index, template_knn_indices


def infer_instance(opts: InferOpts, extractor, repre, grid_points, image_chw: torch.Tensor, mask: torch.Tensor,
                   visual_words_knn_index, template_knn_indices) -> Tuple[List[Dict], Dict[str, float]]:
    """The per-instance block of the reference (scripts/infer.py:467-545), same calls, same order."""
    times: Dict[str, float] = {}
    timer = misc.Timer(enabled=True, cuda_sync=True)
    timer.start()
    extractor_output = extractor(image_chw.unsqueeze(0))
    feature_map_chw = extractor_output["feature_maps"][0]
    times["feat_extract"] = timer.elapsed("Time for feature extraction")
    timer.start()
    query_points = feature_util.filter_points_by_mask(grid_points, mask)
    if query_points.shape[0] > opts.max_num_queries:
        perm = torch.randperm(query_points.shape[0])
        query_points = query_points[perm[: opts.max_num_queries].to(query_points.device)]
    query_features = feature_util.sample_feature_map_at_points(
        feature_map_chw=feature_map_chw, points=query_points,
        image_size=(image_chw.shape[2], image_chw.shape[1])).contiguous()
    times["grid_sample"] = timer.elapsed("Time for grid sample")
    timer.start()
    if query_features.shape[1] != repre.feat_vectors.shape[1] and len(repre.feat_raw_projectors) != 0:
        query_features_proj = projector_util.project_features(
            feat_vectors=query_features, projectors=repre.feat_raw_projectors).contiguous()
    else:
        query_features_proj = query_features
    times["proj"] = timer.elapsed("Time for projection")
    timer.start()
    corresp: List[Dict] = []
    if len(query_points) != 0:
        corresp = corresp_util.establish_correspondences(
            query_points=query_points, query_features=query_features_proj, object_repre=repre,
            template_matching_type=opts.match_template_type, template_knn_indices=template_knn_indices,
            feat_matching_type=opts.match_feat_matching_type, top_n_templates=opts.match_top_n_templates,
            top_k_buddies=opts.match_top_k_buddies, visual_words_knn_index=visual_words_knn_index,
            debug=opts.debug)
    times["corresp"] = timer.elapsed("Time for corresp")
    return corresp, times


def estimate_coarse_poses(opts: InferOpts, corresp: List[Dict], camera_c2w: structs.PinholePlaneCameraModel,
                          seed: int = 0) -> Tuple[List[Dict[str, Any]], int]:
    """Coarse pose per retrieved template and the best one (scripts/infer.py:551-604), same control flow."""
    coarse_poses: List[Dict[str, Any]] = []
    for corresp_id, corresp_curr in enumerate(corresp):
        num_corresp = len(corresp_curr["coord_2d"])
        if num_corresp < 6:
            continue
        ok, R_m2c, t_m2c, inliers, quality = pnp_util.estimate_pose(
            corresp=corresp_curr, camera_c2w=camera_c2w, pnp_type=opts.pnp_type,
            pnp_ransac_iter=opts.pnp_ransac_iter, pnp_inlier_thresh=opts.pnp_inlier_thresh,
            pnp_required_ransac_conf=opts.pnp_required_ransac_conf, pnp_refine_lm=opts.pnp_refine_lm, seed=seed)
        if ok:
            coarse_poses.append({"type": "coarse", "R_m2c": R_m2c, "t_m2c": t_m2c, "corresp_id": corresp_id,
                                 "quality": quality, "inliers": inliers})
    best_quality, best_id = None, 0
    for pose_id, pose in enumerate(coarse_poses):
        if best_quality is None or pose["quality"] > best_quality:
            best_id, best_quality = pose_id, pose["quality"]
    return coarse_poses, best_id


def default_crop_camera(crop_size: Tuple[int, int]) -> structs.PinholePlaneCameraModel:
    """Camera used when crops come without one (synthetic mode): f = 600 px, principal point at the centre."""
    return structs.PinholePlaneCameraModel(crop_size[0], crop_size[1], (600.0, 600.0),
                                           (crop_size[0] / 2.0, crop_size[1] / 2.0))


def make_synthetic_repre(extractor_dim: int, feat_dim: int, templates: int, patches: int, words: int,
                         device: torch.device, seed: int = 0) -> repre_util.FeatureBasedObjectRepre:
    bank = synthetic.make_bank_tensors(templates, patches, feat_dim, num_words=words, seed=seed)
    feat = bank["feat_vectors"].to(device)
    centroids = bank["feat_cluster_centroids"].to(device)
    wk = knn_util.KNN(k=1, metric="l2")
    wk.fit(centroids)
    f2w = wk.search(feat)[1].flatten()
    opts = repre_util.TemplateDescOpts()
    descs, idfs = template_util.calc_tfidf_descriptors(
        feat, f2w, bank["feat_to_template_ids"].to(device), centroids, templates, opts.tfidf_knn_k,
        opts.tfidf_soft_assign, opts.tfidf_soft_sigma_squared)
    projectors = []
    if extractor_dim != feat_dim:
        projectors = [projector_util.projector_from_tensordict(synthetic.make_pca(extractor_dim, feat_dim, seed))]
    return repre_util.FeatureBasedObjectRepre(
        vertices=bank["vertices"], feat_vectors=bank["feat_vectors"],
        feat_to_template_ids=bank["feat_to_template_ids"], feat_to_vertex_ids=bank["feat_to_vertex_ids"],
        feat_to_cluster_ids=f2w.to(torch.int32).cpu(), feat_cluster_centroids=bank["feat_cluster_centroids"],
        feat_cluster_idfs=idfs.cpu(), template_descs=descs.cpu(), template_desc_opts=opts,
        feat_opts=repre_util.FeatureOpts(extractor_name="synthetic"), feat_raw_projectors=projectors)


def infer(opts: InferOpts, repre_dir: Optional[str] = None, crops_path: Optional[str] = None,
          num_synthetic_crops: int = 8, batch: int = 0, output_path: Optional[str] = None,
          synthetic_bank: Tuple[int, int, int, int] = (64, 256, 256, 256)) -> List[Dict[str, Any]]:
    logger = logging.get_logger(level=logging.INFO if opts.debug else logging.WARNING)
    rank, world, local_rank = distributed.init_from_env()
    if not torch.cuda.is_available():
        raise RuntimeError("foundpose_b200 needs a CUDA device (there is no CPU fallback)")
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)

    extractor = feature_util.make_feature_extractor(opts.extractor_name)
    extractor.to(device)

    if repre_dir is not None:
        repre = repre_util.load_object_repre(repre_dir, tensor_device=str(device)) if rank == 0 else None
    else:
        t, p, d, w = synthetic_bank
        repre = make_synthetic_repre(extractor.arch.embed_dim, d, t, p, w, device) if rank == 0 else None
    repre = distributed.broadcast_object_repre(repre, src=0, device=device)

    # Crop cameras (fx, fy, cx, cy per crop).  Real crops must bring the intrinsics of their virtual crop camera
    # (the reference builds it per instance with misc.construct_crop_camera, scripts/infer.py:411-440): without them no
    # coarse pose is emitted.  Only the synthetic mode falls back to a made-up camera.
    crop_intrinsics: Optional[torch.Tensor] = None
    if crops_path is not None:
        blob = torch.load(crops_path, map_location="cpu")
        images, masks = blob["images"].float(), blob["masks"].bool()
        if "intrinsics" in blob:
            crop_intrinsics = torch.as_tensor(blob["intrinsics"], dtype=torch.float64).reshape(-1, 4)
            if crop_intrinsics.shape[0] != images.shape[0]:
                raise ValueError("crops blob: `intrinsics` needs one (fx, fy, cx, cy) row per crop")
        else:
            logger.warning("The crops blob has no `intrinsics` [P,4]: correspondences only, no coarse poses.")
    else:
        images = synthetic.make_crops(num_synthetic_crops, opts.crop_size, seed=1)
        masks = synthetic.make_masks(num_synthetic_crops, opts.crop_size, seed=2)
        cam = default_crop_camera(opts.crop_size)
        crop_intrinsics = torch.from_numpy(pnp_util.get_intrinsics_vector(cam)).reshape(1, 4).repeat(images.shape[0], 1)
    start, end = distributed.shard_range(images.shape[0], rank, world)
    images, masks = images[start:end].to(device), masks[start:end].to(device)
    if crop_intrinsics is not None:
        crop_intrinsics = crop_intrinsics[start:end].to(device).contiguous()

    grid_points = feature_util.generate_grid_points(grid_size=opts.crop_size, cell_size=opts.grid_cell_size).to(device)
    index, visual_words_knn_index, template_knn_indices = build_indices(repre, opts, device)
    logging.log_heading(logger, f"Object representation: {index.num_templates} templates, "
                                f"{index.bank16.shape[0]} features, vertices: {len(repre.vertices)}")
    results: List[Dict[str, Any]] = []
    if batch > 0:
        grid_n = grid_points.shape[0]
        if opts.max_num_queries < grid_n:
            # the reference sub-samples the query points per instance with torch.randperm (scripts/infer.py:483-485);
            # the batched path keeps a fixed stride of grid points per crop
            raise ValueError(f"max_num_queries={opts.max_num_queries} < {grid_n} grid points: use the per-crop path "
                             "(batch=0), which sub-samples the query points like the reference")
        pipe = pipeline.CropBatchPipeline(extractor, index, repre.feat_raw_projectors, batch,
                                          crop_size=opts.crop_size, grid_cell_size=opts.grid_cell_size,
                                          top_n_templates=opts.match_top_n_templates,
                                          top_k_buddies=opts.match_top_k_buddies)
        n = images.shape[0]
        for s in range(0, n, batch):
            e = min(n, s + batch)
            img = images[s:e]
            msk = masks[s:e].to(torch.uint8)
            if e - s < batch:   # pad the last batch
                img = torch.cat([img, img[:1].expand(batch - (e - s), -1, -1, -1)])
                msk = torch.cat([msk, torch.zeros((batch - (e - s),) + msk.shape[1:], dtype=torch.uint8, device=device)])
            out = pipe.run(img.contiguous(), msk.contiguous())
            # Coarse poses of all (crop, template) pairs of the batch in one launch.
            topn, kk = out.count.shape[1], out.coord_2d.shape[2]
            best = poses = None
            if crop_intrinsics is not None:
                intr = crop_intrinsics[s:e]
                if e - s < batch:
                    intr = torch.cat([intr, intr[:1].expand(batch - (e - s), 4)])
                poses = pnp_util.estimate_poses_batched(
                    out.coord_2d.reshape(batch * topn, kk, 2), out.coord_3d.reshape(batch * topn, kk, 3),
                    out.count.reshape(-1), intr.repeat_interleave(topn, dim=0).contiguous(), opts.pnp_ransac_iter,
                    opts.pnp_inlier_thresh, opts.pnp_required_ransac_conf, problem_offset=(start + s) * topn)
                best = pnp_util.select_best_poses(poses["success"], poses["num_inliers"], topn).cpu()
                poses = {k: v.cpu() for k, v in poses.items()}
            for b in range(e - s):
                corresp = pipeline.outputs_to_corresp_list(out, b)
                j = int(best[b]) if best is not None else -1
                results.append({"crop_id": start + s + b, "corresp": [
                    {k: v.cpu().clone() for k, v in c.items()} for c in corresp],
                    "best_coarse_pose": None if j < 0 else {
                        "corresp_id": j, "R_m2c": poses["R"][b * topn + j].numpy(),
                        "t_m2c": poses["t"][b * topn + j].numpy().reshape(3, 1),
                        "quality": float(poses["num_inliers"][b * topn + j])}})
    else:
        for i in range(images.shape[0]):
            corresp, times = infer_instance(opts, extractor, repre, grid_points, images[i], masks[i],
                                            visual_words_knn_index, template_knn_indices)
            logger.info(f"Number of corresp: {[len(c['coord_2d']) for c in corresp]}")
            coarse_poses, best_id = [], 0
            if crop_intrinsics is not None:
                fx, fy, cx, cy = [float(v) for v in crop_intrinsics[i].tolist()]
                camera_c2w = structs.PinholePlaneCameraModel(opts.crop_size[0], opts.crop_size[1], (fx, fy), (cx, cy))
                coarse_poses, best_id = estimate_coarse_poses(opts, corresp, camera_c2w)
            results.append({"crop_id": start + i, "time": times,
                            "corresp": [{k: v.cpu() for k, v in c.items()} for c in corresp],
                            "coarse_poses": coarse_poses,
                            "best_coarse_pose": coarse_poses[best_id] if coarse_poses else None})
    if output_path is not None:
        suffix = f".rank{rank}" if world > 1 else ""
        torch.save(results, output_path + suffix)
    return results


def main() -> None:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--opts-path", type=str, default=None)
    ap.add_argument("--repre-dir", type=str, default=None)
    ap.add_argument("--crops", type=str, default=None)
    ap.add_argument("--synthetic", action="store_true")
    ap.add_argument("--num-crops", type=int, default=8)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--output", type=str, default=None)
    for name in ("extractor_name", "grid_cell_size", "match_top_n_templates", "match_top_k_buddies"):
        ap.add_argument(f"--{name.replace('_', '-')}", default=None,
                        type={"extractor_name": str, "grid_cell_size": float}.get(name, int))
    args = ap.parse_args()
    overrides = {k: getattr(args, k) for k in ("extractor_name", "grid_cell_size", "match_top_n_templates",
                                               "match_top_k_buddies")}
    opts = load_opts(args.opts_path, overrides)
    if args.repre_dir is None and not args.synthetic:
        ap.error("give --repre-dir <dir with repre.pth> or --synthetic")
    results = infer(opts, repre_dir=args.repre_dir, crops_path=args.crops, num_synthetic_crops=args.num_crops,
                    batch=args.batch, output_path=args.output)
    for r in results[:4]:
        ids = [int(c["template_id"]) for c in r["corresp"]]
        print(f"crop {r['crop_id']}: templates {ids}, corresp {[len(c['coord_2d']) for c in r['corresp']]}")


if __name__ == "__main__":
    main()
