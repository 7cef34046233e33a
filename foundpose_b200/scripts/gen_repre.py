#!/usr/bin/env python3
"""Offline object-representation build on the GPU - mirror of the reference's scripts/gen_repre.py.

The reference renders templates with pyrender / BOP tooling (scripts/gen_templates.py, out of scope,
SURVEY.md §2) and then, per object (scripts/gen_repre.py:66-377): extracts DINOv2 features of every
template, registers them in 3D, fits a PCA, clusters the projected features into visual words, computes
tf-idf template descriptors and saves `repre.pth`.  This module runs that second half from template data
you provide (image, depth, mask, camera and pose per template), with the same options (`GenRepreOpts`),
the same calls in the same order and the same output container:

  * feature extraction: ONE batched extractor call per chunk of templates (the reference: one call per
    template), `feature_util.get_visual_features_registered_in_3d` for the 3D registration;
  * PCA fit: covariance on the tcgen05 GEMM + eigendecomposition (`projector_util.PCAProjector.fit`, scikit-learn's
    `covariance_eigh` conventions), transform on the GPU;
  * k-means: `cluster_util.kmeans` (tcgen05 assignment + fixed-point update kernels);
  * tf-idf descriptors: `template_util.calc_tfidf_descriptors`;
  * `repre_util.save_object_repre`: the reference's `repre.pth` layout.
"""

from __future__ import annotations

import os
import sys
from typing import Any, Dict, List, NamedTuple, Optional, Sequence

import numpy as np
import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from foundpose_b200.utils import (cluster_util, feature_util, logging, misc, projector_util,  # noqa: E402
                                  repre_util, structs, template_util)


class GenRepreOpts(NamedTuple):
    """Options that can be specified via the command line (reference scripts/gen_repre.py:37-64)."""

    version: str = "v1"
    templates_version: str = "v1"
    object_dataset: str = "lmo"
    object_lids: Optional[List[int]] = None

    # Feature extraction options.
    extractor_name: str = "dinov2_vits14_reg"
    grid_cell_size: float = 14.0

    # Feature PCA options.
    apply_pca: bool = True
    pca_components: int = 256
    pca_whiten: bool = False
    pca_max_samples_for_fitting: int = 100000

    # Feature clustering options.
    cluster_features: bool = True
    cluster_num: int = 2048

    # Template descriptor options.
    template_desc_opts: Optional[repre_util.TemplateDescOpts] = None

    # Other options.
    overwrite: bool = True
    debug: bool = True


class TemplateSample(NamedTuple):
    """What the reference reads per template from the rendered template folder (gen_repre.py:107-150)."""

    image_chw: torch.Tensor              # fp32 [3, H, W] in [0, 1]
    depth_image_hw: torch.Tensor         # fp32 [H, W], same unit as the model (mm)
    object_mask: torch.Tensor            # [H, W], non-zero = object
    camera: structs.PinholePlaneCameraModel   # intrinsics + T_world_from_eye
    T_world_from_model: np.ndarray       # 4x4 object pose


def generate_raw_repre(opts: GenRepreOpts, templates: Sequence[TemplateSample], extractor: torch.nn.Module,
                       device: torch.device, extract_batch: int = 32) -> repre_util.FeatureBasedObjectRepre:
    """Features of all templates registered in 3D (reference generate_raw_repre, gen_repre.py:66-217)."""
    feat_vectors_list, feat_to_vertex_ids_list, vertices_in_model_list = [], [], []
    feat_to_template_ids_list, templates_list, cameras_list = [], [], []
    for start in range(0, len(templates), extract_batch):
        chunk = templates[start:start + extract_batch]
        images = torch.stack([t.image_chw.to(torch.float32) for t in chunk]).to(device)
        feature_maps = extractor(images)["feature_maps"]                       # one launch sequence per chunk
        for j, t in enumerate(chunk):
            template_id = start + j
            T_world_from_model = torch.as_tensor(np.asarray(t.T_world_from_model), dtype=torch.float32, device=device)
            T_model_from_world = torch.linalg.inv(T_world_from_model)
            T_world_from_camera = torch.as_tensor(t.camera.T_world_from_eye, dtype=torch.float32, device=device)
            T_model_from_camera = torch.matmul(T_model_from_world, T_world_from_camera)
            feat_vectors, feat_to_vertex_ids, vertices_in_model = feature_util.get_visual_features_registered_in_3d(
                image_chw=images[j], depth_image_hw=t.depth_image_hw.to(device, torch.float32),
                object_mask=t.object_mask.to(device, torch.float32), camera=t.camera,
                T_model_from_camera=T_model_from_camera, extractor=extractor, grid_cell_size=opts.grid_cell_size,
                feature_map_chw=feature_maps[j])
            feat_vectors_list.append(feat_vectors)
            feat_to_vertex_ids_list.append(feat_to_vertex_ids)
            vertices_in_model_list.append(vertices_in_model)
            feat_to_template_ids_list.append(
                template_id * torch.ones(feat_vectors.shape[0], dtype=torch.int32, device=device))
            templates_list.append((images[j] * 255).to(torch.uint8))
            cameras_list.append({"f": torch.as_tensor(t.camera.f), "c": torch.as_tensor(t.camera.c),
                                 "width": t.camera.width, "height": t.camera.height,
                                 "T_world_from_eye": torch.linalg.inv(T_model_from_camera).cpu()})
    return repre_util.FeatureBasedObjectRepre(
        vertices=torch.cat(vertices_in_model_list), feat_vectors=torch.cat(feat_vectors_list),
        feat_opts=repre_util.FeatureOpts(extractor_name=opts.extractor_name),
        feat_to_vertex_ids=torch.cat(feat_to_vertex_ids_list),
        feat_to_template_ids=torch.cat(feat_to_template_ids_list), templates=torch.stack(templates_list),
        template_cameras_cam_from_model=cameras_list)


def generate_repre(opts: GenRepreOpts, templates: Sequence[TemplateSample], device: str = "cuda",
                   extractor: Optional[torch.nn.Module] = None, output_dir: Optional[str] = None
                   ) -> repre_util.FeatureBasedObjectRepre:
    """The reference's generate_repre (gen_repre.py:220-377) from in-memory templates."""
    logger = logging.get_logger(level=logging.INFO if opts.debug else logging.WARNING)
    dev = torch.device(device)
    if output_dir is not None:
        if os.path.exists(os.path.join(output_dir, "repre.pth")) and not opts.overwrite:
            raise ValueError(f"Output directory already exists: {output_dir}")
        os.makedirs(output_dir, exist_ok=True)
    timer = misc.Timer(enabled=opts.debug)
    timer.start()
    if extractor is None:
        extractor = feature_util.make_feature_extractor(opts.extractor_name)
    extractor.to(dev)
    repre = generate_raw_repre(opts, templates, extractor, dev)
    feat_vectors = repre.feat_vectors
    assert feat_vectors is not None
    timer.elapsed("Time for generating raw representation")

    if opts.apply_pca:
        timer.start()
        pca_projector = projector_util.PCAProjector(n_components=opts.pca_components, whiten=opts.pca_whiten)
        pca_projector.fit(feat_vectors, max_samples=opts.pca_max_samples_for_fitting)
        repre.feat_raw_projectors.append(pca_projector)
        feat_vectors = pca_projector.transform(feat_vectors).contiguous()
        timer.elapsed("Time for PCA")

    if opts.cluster_features:
        timer.start()
        centroids, cluster_ids, _ = cluster_util.kmeans(samples=feat_vectors, num_centroids=opts.cluster_num,
                                                        verbose=opts.debug)
        repre.feat_cluster_centroids = centroids
        repre.feat_to_cluster_ids = cluster_ids
        _, unique_counts = torch.unique(cluster_ids, return_counts=True)
        timer.elapsed("Time for feature clustering")
        logging.log_heading(logger, f"{feat_vectors.shape[0]} feature vectors were clustered into {len(centroids)} "
                                    f"clusters with {unique_counts.min()} to {unique_counts.max()} elements.")

    if opts.template_desc_opts is not None:
        timer.start()
        repre.template_desc_opts = opts.template_desc_opts
        if opts.template_desc_opts.desc_type == "tfidf":
            assert repre.feat_cluster_centroids is not None and repre.feat_to_cluster_ids is not None
            assert repre.feat_to_template_ids is not None and repre.templates is not None
            repre.template_descs, repre.feat_cluster_idfs = template_util.calc_tfidf_descriptors(
                feat_vectors, repre.feat_to_cluster_ids, repre.feat_to_template_ids, repre.feat_cluster_centroids,
                len(repre.templates), opts.template_desc_opts.tfidf_knn_k, opts.template_desc_opts.tfidf_soft_assign,
                opts.template_desc_opts.tfidf_soft_sigma_squared)
        else:
            raise ValueError(f"Unknown template descriptor type: {opts.template_desc_opts.desc_type}")
        timer.elapsed("Time for generating template descriptors")

    if len(repre.feat_raw_projectors) and isinstance(repre.feat_raw_projectors[0], projector_util.PCAProjector):
        repre.feat_vis_projectors = [repre.feat_raw_projectors[0]]
    repre.feat_vectors = feat_vectors
    if output_dir is not None:
        repre_util.save_object_repre(repre, output_dir)
    return repre
