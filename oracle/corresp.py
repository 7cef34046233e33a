"""Oracle: cyclic-buddies 2D-3D correspondences.  TEST INFRASTRUCTURE ONLY.

Restates utils/corresp_util.py:34-169.  Where the reference's `torch.topk(-cycle_dists)` leaves
the order of exact ties implementation-defined the oracle uses the canonical
order: ascending cyclic distance, ties by ascending query index.
"""

from __future__ import annotations

from typing import Dict, List, Tuple

import torch

from . import knn as oknn
from . import template as otemplate


def cyclic_buddies_matching(query_points: torch.Tensor, query_features: torch.Tensor,
                            object_features: torch.Tensor, top_k: int
                            ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """utils/corresp_util.py:34-70."""
    query2obj = oknn.knn_l2(query_features, object_features, 1)[1].flatten()
    obj2query = oknn.knn_l2(object_features, query_features, 1)[1].flatten()
    return cyclic_buddies_from_ids(query_points, query2obj, obj2query, top_k)


def cyclic_buddies_from_ids(query_points: torch.Tensor, query2obj: torch.Tensor, obj2query: torch.Tensor, top_k: int
                            ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """utils/corresp_util.py:50-70: the part after the two 1-NN searches (cycle, 2D distance, top-k, scores)."""
    u1 = query_points
    cycle_ids = obj2query[query2obj]
    u2 = query_points[cycle_ids]
    cycle_dists = torch.linalg.norm(u1 - u2, axis=1)
    top_k = min(top_k, query_points.shape[0])
    query_bb_ids = torch.sort(cycle_dists, stable=True).indices[:top_k]
    bb_dists = cycle_dists[query_bb_ids]
    bb_scores = torch.as_tensor(1.0 - (bb_dists / bb_dists.max()))
    object_bb_ids = query2obj[query_bb_ids]
    return query_bb_ids, object_bb_ids, bb_dists, bb_scores


def establish_correspondences(query_points: torch.Tensor, query_features: torch.Tensor, repre: Dict,
                              top_n_templates: int, top_k_buddies: int, knn_k: int = 3,
                              soft_assign: bool = False, soft_sigma_squared: float = 10.0) -> List[Dict]:
    """utils/corresp_util.py:73-169 on a dict of bank tensors (fields of repre_util.py:34-83)."""
    template_ids, template_scores, _, _ = otemplate.tfidf_matching(
        query_features, repre["feat_cluster_centroids"], repre["feat_cluster_idfs"],
        repre["template_descs"], top_n_templates, knn_k, "l2", soft_assign, soft_sigma_squared)
    corresps = []
    for counter, template_id in enumerate(template_ids):
        tpl_feat_ids = torch.nonzero(repre["feat_to_template_ids"] == template_id).flatten()
        q_ids, o_ids, dists, scores = cyclic_buddies_matching(
            query_points, query_features, repre["feat_vectors"][tpl_feat_ids], top_k_buddies)
        obj_feat_ids = tpl_feat_ids[o_ids]
        corresps.append({
            "template_id": template_id,
            "template_score": template_scores[counter],
            "coord_2d": query_points[q_ids],
            "coord_2d_ids": q_ids,
            "coord_3d": repre["vertices"][obj_feat_ids],
            "coord_conf": scores,
            "nn_vertex_ids": obj_feat_ids,
            "nn_dists": dists,
            "nn_indices": obj_feat_ids,
        })
    return corresps
