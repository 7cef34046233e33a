"""Cyclic-buddies 2D-3D correspondences - mirror of the reference's utils/corresp_util.py."""

from typing import Any, Dict, List, Optional, Tuple

import torch

from foundpose_b200 import _native, pipeline
from foundpose_b200.utils import knn_util, logging, misc, repre_util, template_util

logger: logging.Logger = logging.get_logger()

_MAX_ENGINES = 8   # RetrievalEngine instances cached per ObjectIndex (one per 128-row bucket of query counts)


def convert_px_indices_to_im_coords(px_indices: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    """Pixel index (i, j) -> image coordinates (i + 0.5, j + 0.5), optionally scaled."""
    return scale * (px_indices.float() + 0.5)


def cyclic_buddies_matching(query_points: torch.Tensor, query_features: torch.Tensor,
                            query_knn_index: knn_util.KNN, object_features: torch.Tensor,
                            object_knn_index: knn_util.KNN, top_k: int, debug: bool
                            ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """Best buddies via cyclic distance (reference :34-70).

    Ties in the cyclic distance (frequent: distances between grid points) are returned in the
    canonical order (distance, then query index); the reference's torch.topk leaves it undefined.
    """
    if not query_points.is_cuda:
        raise ValueError("foundpose_b200 runs on CUDA tensors only (no CPU fallback)")
    dev = query_points.device
    query2obj = object_knn_index.search(query_features)[1].to(dev).contiguous()
    obj2query = query_knn_index.search(object_features)[1].to(dev).contiguous()
    n = query_points.shape[0]
    p = obj2query.shape[0]
    k = min(top_k, n)
    i32, i64, f32 = torch.int32, torch.int64, torch.float32
    pts = query_points.to(f32).contiguous()
    out_q = torch.empty((max(top_k, 1),), dtype=i64, device=dev)
    out_v = torch.empty_like(out_q)
    out_d = torch.empty((max(top_k, 1),), dtype=f32, device=dev)
    out_s = torch.empty_like(out_d)
    out_c2 = torch.empty((max(top_k, 1), 2), dtype=f32, device=dev)
    out_c3 = torch.empty((max(top_k, 1), 3), dtype=f32, device=dev)
    out_n = torch.zeros((1,), dtype=i32, device=dev)
    if n > 0:
        _native.cyclic_buddies(
            pts, torch.zeros(1, dtype=i32, device=dev), torch.full((1,), n, dtype=i32, device=dev),
            query2obj, obj2query, torch.zeros(1, dtype=i64, device=dev), 1,
            torch.tensor([0, p], dtype=i32, device=dev), None, torch.zeros((max(p, 1), 3), dtype=f32, device=dev),
            n, max(p, 1), max(top_k, 1), out_q, out_v, out_d, out_s, out_c2, out_c3, out_n)
    return out_q[:k], out_v[:k], out_d[:k], out_s[:k]


def establish_correspondences(query_points: torch.Tensor, query_features: torch.Tensor,
                              object_repre: repre_util.FeatureBasedObjectRepre, template_matching_type: str,
                              feat_matching_type: str, top_n_templates: int, top_k_buddies: int,
                              visual_words_knn_index: Optional[knn_util.KNN] = None,
                              template_knn_indices: Optional[List[knn_util.KNN]] = None,
                              debug: bool = False) -> List[Dict]:
    """Establishes 2D-3D correspondences by matching image and object features (reference :73-169).

    The whole stage (visual-word k-NN, tf-idf, cosine retrieval, 2 x top_n 1-NN searches, cyclic
    distances, top-k, gathers) runs as one launch sequence of pipeline.RetrievalEngine on the packed
    ObjectIndex of `object_repre`; `visual_words_knn_index` / `template_knn_indices` are accepted for
    signature compatibility (their contents are the same bank rows the ObjectIndex holds).
    """
    if template_matching_type != "tfidf":
        raise ValueError(f"Unknown matching type '{template_matching_type}'.")
    if feat_matching_type != "cyclic_buddies":
        raise ValueError(f"Unknown feature matching type ({feat_matching_type}).")
    if object_repre.template_desc_opts is None or object_repre.template_desc_opts.desc_type != "tfidf":
        raise ValueError("Template descriptors need to be tfidf.")
    if not query_features.is_cuda:
        raise ValueError("foundpose_b200 runs on CUDA tensors only (no CPU fallback)")
    timer = misc.Timer(enabled=debug)
    timer.start()
    dev = query_features.device
    index = pipeline.get_object_index(object_repre, dev)
    n = query_points.shape[0]
    if n == 0:
        return []
    knn_k = visual_words_knn_index.k if visual_words_knn_index is not None else None
    # One engine (device buffers sized for `stride` query rows) serves every crop whose point count falls in the
    # same 128-row bucket; the actual count goes in as q_count.  At most _MAX_ENGINES stay cached per object.
    stride = (n + 127) // 128 * 128
    key = (stride, top_n_templates, top_k_buddies, knn_k)
    engines = index.__dict__.setdefault("_engines", {})
    engine = engines.pop(key, None)
    if engine is None:
        engine = pipeline.RetrievalEngine(index, 1, stride, top_n_templates, top_k_buddies, knn_k)
        engine.stage_feat = torch.zeros((stride, index.dim_padded), dtype=torch.float32, device=dev)
        engine.stage_pts = torch.zeros((1, stride, 2), dtype=torch.float32, device=dev)
        engine.stage_feat16 = torch.zeros((stride, index.dim_padded), dtype=torch.float16, device=dev)
    engines[key] = engine                                   # most recently used last
    while len(engines) > _MAX_ENGINES:
        engines.pop(next(iter(engines)))
    engine.stage_feat[:n, : query_features.shape[1]] = query_features.to(torch.float32)
    engine.stage_pts[0, :n] = query_points.to(dev, torch.float32)
    _native.convert_rows_f16(engine.stage_feat, out=engine.stage_feat16)
    count = torch.full((1,), n, dtype=torch.int32, device=dev)
    out = engine.match(engine.stage_feat16, engine.stage_pts, count)
    corresps = pipeline.outputs_to_corresp_list(out, 0, debug=debug)
    # The engine's buffers are reused by the next call: hand out copies.
    corresps = [{k: v.clone() for k, v in c.items()} for c in corresps]
    logger.info(f"Matched templates: {[int(c['template_id']) for c in corresps]}")
    timer.elapsed("Time for establishing corresp")
    return corresps
