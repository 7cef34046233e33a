"""Oracle: tf-idf bag-of-visual-words template retrieval.  TEST INFRASTRUCTURE ONLY.

Restates utils/template_util.py:13-176 with the same torch ops (CPU, fp32).
"""

from __future__ import annotations

from typing import Tuple

import torch

from . import knn as oknn


def find_nearest_object_features(query_features: torch.Tensor, centroids: torch.Tensor, k: int,
                                 metric: str = "l2") -> Tuple[torch.Tensor, torch.Tensor]:
    """utils/template_util.py:13-29: k-NN vs the visual words, sqrt of faiss's squared distances.

    Returns (ids [Nq,k] int64, dists [Nq,k] fp32) - note the order.
    """
    if metric == "l2":
        d, i = oknn.knn_l2(query_features, centroids, k)
    else:
        d, i = oknn.knn_cosine(query_features, centroids, k)
    return i, torch.sqrt(d)


def calc_tfidf(feature_word_ids: torch.Tensor, feature_word_dists: torch.Tensor, word_idfs: torch.Tensor,
               soft_assignment: bool = True, soft_sigma_squared: float = 100.0) -> torch.Tensor:
    """utils/template_util.py:31-71."""
    if soft_assignment:
        w = torch.exp(-torch.square(feature_word_dists) / (2.0 * soft_sigma_squared))
    else:
        w = torch.ones_like(feature_word_dists)
    w = torch.nn.functional.normalize(w, p=2, dim=1).reshape(-1)
    tf = w / feature_word_ids.shape[0]
    flat = feature_word_ids.reshape(-1)
    idf = word_idfs[flat]
    tfidf = torch.multiply(tf, idf)
    num_words = word_idfs.shape[0]
    return torch.zeros(num_words, dtype=w.dtype).scatter_add_(dim=0, index=flat.to(torch.int64), src=tfidf)


def calc_tfidf_descriptors(feat_vectors: torch.Tensor, feat_to_word_ids: torch.Tensor,
                           feat_to_template_ids: torch.Tensor, feat_words: torch.Tensor,
                           num_templates: int, tfidf_knn_k: int, tfidf_soft_assign: bool,
                           tfidf_soft_sigma_squared: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """utils/template_util.py:74-123 (offline; defines template_descs / idfs of the bank).

    Note: unlike the query side, squared distances are fed to calc_tfidf here.
    """
    occ = torch.zeros(len(feat_words), dtype=torch.int64)
    for t in range(num_templates):
        m = feat_to_template_ids == t
        occ[torch.unique(feat_to_word_ids[m])] += 1
    idfs = torch.log(torch.as_tensor(float(num_templates)) / occ.to(torch.float32))
    descs = []
    for t in range(num_templates):
        m = feat_to_template_ids == t
        d, i = oknn.knn_l2(feat_vectors[m], feat_words, tfidf_knn_k)
        descs.append(calc_tfidf(i, d, idfs, tfidf_soft_assign, tfidf_soft_sigma_squared))
    return torch.stack(descs, dim=0), idfs


def calc_tfidf_descriptors_vectorised(feat_vectors: torch.Tensor, feat_to_word_ids: torch.Tensor,
                                      feat_to_template_ids: torch.Tensor, feat_words: torch.Tensor,
                                      num_templates: int, tfidf_knn_k: int, chunk: int = 1 << 18
                                      ) -> Tuple[torch.Tensor, torch.Tensor]:
    """`calc_tfidf_descriptors` (hard assignment) without the per-template Python loops.

    Same arithmetic as utils/template_util.py:74-123 with tfidf_soft_assign=False - idf_i = log(N / N_i) from
    the nearest-word assignment, every feature adds (1/sqrt(k)) / n_t * idf to each of its k nearest words of
    its template's descriptor - expressed as chunked matrix products and one scatter_add, so that the 10 k-template
    bank of BASELINE configs[2] can be set up in seconds (on whatever device the tensors live on).  Only used to
    PREPARE benchmark inputs of the reference arm; checked against the loop version in tests/test_oracle_golden.py.
    """
    dev = feat_vectors.device
    num_words = feat_words.shape[0]
    tpl = feat_to_template_ids.to(torch.int64)
    pair = torch.unique(tpl * num_words + feat_to_word_ids.to(torch.int64))
    occ = torch.bincount(pair % num_words, minlength=num_words)
    idfs = torch.log(torch.as_tensor(float(num_templates), device=dev) / occ.to(torch.float32))
    counts = torch.bincount(tpl, minlength=num_templates).to(torch.float32)
    words = feat_words.to(torch.float32)
    wn = (words * words).sum(dim=1).unsqueeze(0)
    descs = torch.zeros(num_templates * num_words, dtype=torch.float32, device=dev)
    w = 1.0 / (tfidf_knn_k ** 0.5)
    for s in range(0, feat_vectors.shape[0], chunk):
        x = feat_vectors[s:s + chunk].to(torch.float32)
        d = (x * x).sum(dim=1, keepdim=True) + wn - 2.0 * (x @ words.t())
        ids = torch.topk(d, tfidf_knn_k, dim=1, largest=False).indices            # [n, k]
        t = tpl[s:s + chunk]
        val = (w / counts[t]).unsqueeze(1) * idfs[ids]
        descs.scatter_add_(0, (t.unsqueeze(1) * num_words + ids).reshape(-1), val.reshape(-1))
    return descs.reshape(num_templates, num_words), idfs


def tfidf_matching(query_features: torch.Tensor, centroids: torch.Tensor, idfs: torch.Tensor,
                   template_descs: torch.Tensor, top_n_templates: int, knn_k: int = 3,
                   knn_metric: str = "l2", soft_assign: bool = False,
                   soft_sigma_squared: float = 10.0):
    """utils/template_util.py:126-176.

    Returns (template_ids [n] int64, template_scores [n] fp32, query_tfidf [W], cos_sims [T]).
    Ties in the final top-k are broken by ascending template id (canonical rule).
    """
    word_ids, word_dists = find_nearest_object_features(query_features, centroids, knn_k, knn_metric)
    query_tfidf = calc_tfidf(word_ids, word_dists, idfs, soft_assign, soft_sigma_squared)
    num_templates = template_descs.shape[0]
    cos = torch.nn.functional.cosine_similarity(template_descs, query_tfidf.tile(num_templates, 1))
    order = torch.sort(-cos, stable=True)
    ids = order.indices[:top_n_templates]
    return ids, cos[ids], query_tfidf, cos
