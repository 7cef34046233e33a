"""tf-idf bag-of-visual-words template retrieval - mirror of the reference's utils/template_util.py."""

from typing import Optional, Tuple

import torch

from foundpose_b200 import _native, pipeline
from foundpose_b200.utils import knn_util, logging, misc, repre_util

logger: logging.Logger = logging.get_logger()


def find_nearest_object_features(query_features: torch.Tensor, knn_index: knn_util.KNN
                                 ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Nearest reference features of each query; returns (ids, sqrt'ed distances) - note the order."""
    nn_dists, nn_ids = knn_index.search(query_features)
    knn_k = nn_dists.shape[1]
    nn_dists = nn_dists[:, :knn_k]
    nn_ids = nn_ids[:, :knn_k]
    # The distances returned by the index are squared (reference :26-27).
    nn_dists = torch.sqrt(nn_dists)
    return nn_ids, nn_dists


def calc_tfidf(feature_word_ids: torch.Tensor, feature_word_dists: torch.Tensor, word_idfs: torch.Tensor,
               soft_assignment: bool = True, soft_sigma_squared: float = 100.0) -> torch.Tensor:
    """tf-idf descriptor of one image from its (feature -> k visual words) assignment (reference :31-71)."""
    if not feature_word_ids.is_cuda:
        raise ValueError("foundpose_b200 runs on CUDA tensors only (no CPU fallback)")
    dev = feature_word_ids.device
    n = feature_word_ids.shape[0]
    ids = feature_word_ids.to(torch.int64).contiguous()
    dists = feature_word_dists.to(dev, torch.float32).contiguous()
    idfs = word_idfs.to(dev, torch.float32).contiguous()
    out = torch.empty((1, idfs.shape[0]), dtype=torch.float32, device=dev)
    start = torch.zeros(1, dtype=torch.int32, device=dev)
    count = torch.full((1,), n, dtype=torch.int32, device=dev)
    _native.calc_tfidf(ids, dists, start, count, idfs, soft_assignment, soft_sigma_squared, False, out)
    return out[0]


def calc_tfidf_descriptors(feat_vectors: torch.Tensor, feat_to_word_ids: torch.Tensor,
                           feat_to_template_ids: torch.Tensor, feat_words: torch.Tensor, num_templates: int,
                           tfidf_knn_k: int, tfidf_soft_assign: bool, tfidf_soft_sigma_squared: float
                           ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-template tf-idf descriptors and idfs (reference :74-123; offline, defines the bank).

    One k-NN over all bank rows and one batched calc_tfidf over the templates' row runs replace the
    reference's per-template loop.  As in the reference, SQUARED distances go into calc_tfidf here
    (SURVEY.md S9).
    """
    dev = feat_vectors.device if feat_vectors.is_cuda else torch.device("cuda", torch.cuda.current_device())
    feat = feat_vectors.to(dev, torch.float32)
    tpl = feat_to_template_ids.to(dev).to(torch.int64)
    words = feat_to_word_ids.to(dev).to(torch.int64)
    num_words = len(feat_words)
    # idf = log(N / N_i): N_i = number of templates in which word i occurs.
    pair = torch.unique(tpl * num_words + words)
    occ = torch.bincount(pair % num_words, minlength=num_words)
    word_idfs = torch.log(torch.as_tensor(float(num_templates), device=dev) / occ.to(torch.float32))

    order = torch.sort(tpl, stable=True).indices
    counts = torch.bincount(tpl, minlength=num_templates)[:num_templates]
    starts = torch.cumsum(counts, 0) - counts
    index = knn_util.KNN(k=tfidf_knn_k, metric="l2")
    index.fit(feat_words.to(dev))
    word_dists, word_ids = index.search(feat[order].contiguous())
    out = torch.empty((num_templates, num_words), dtype=torch.float32, device=dev)
    _native.calc_tfidf(word_ids.contiguous(), word_dists.contiguous(), starts.to(torch.int32).contiguous(),
                       counts.to(torch.int32).contiguous(), word_idfs.contiguous(), tfidf_soft_assign,
                       tfidf_soft_sigma_squared, False, out)
    return out.to(feat_vectors.device), word_idfs.to(feat_vectors.device)


def tfidf_matching(query_features: torch.Tensor, object_repre: repre_util.FeatureBasedObjectRepre,
                   top_n_templates: int, visual_words_knn_index: knn_util.KNN, debug: bool = False
                   ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Top-N templates by cosine similarity of tf-idf descriptors (reference :126-176)."""
    if object_repre.template_desc_opts is None or object_repre.template_desc_opts.desc_type != "tfidf":
        raise ValueError("Template descriptors need to be tfidf.")
    if not query_features.is_cuda:
        raise ValueError("foundpose_b200 runs on CUDA tensors only (no CPU fallback)")
    timer = misc.Timer(enabled=debug)
    timer.start()
    word_ids, word_dists = find_nearest_object_features(query_features=query_features,
                                                        knn_index=visual_words_knn_index)
    timer.elapsed("Time for KNN search")
    assert object_repre.feat_cluster_idfs is not None
    opts = object_repre.template_desc_opts
    query_tfidf = calc_tfidf(feature_word_ids=word_ids, feature_word_dists=word_dists,
                             word_idfs=object_repre.feat_cluster_idfs,
                             soft_assignment=opts.tfidf_soft_assign,
                             soft_sigma_squared=opts.tfidf_soft_sigma_squared)
    assert object_repre.template_descs is not None
    index = pipeline.get_object_index(object_repre, query_features.device)
    cos = torch.empty((1, index.num_templates), dtype=torch.float32, device=query_features.device)
    _native.bow_scores(index.template_descs, index.desc_norm, query_tfidf.reshape(1, -1).contiguous(), cos)
    k = min(top_n_templates, index.num_templates)
    if top_n_templates > index.num_templates:
        raise RuntimeError("selected index k out of range")
    template_scores = torch.empty((1, k), dtype=torch.float32, device=cos.device)
    template_ids = torch.empty((1, k), dtype=torch.int64, device=cos.device)
    _native.topk_rows(cos, k, template_scores, template_ids)
    return template_ids[0], template_scores[0]


def template_matching(query_features: torch.Tensor, object_repre: repre_util.FeatureBasedObjectRepre,
                      top_n_templates: int, matching_type: str,
                      visual_words_knn_index: Optional[knn_util.KNN] = None
                      ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Retrieves N most similar templates to the query image."""
    if matching_type == "tfidf":
        assert visual_words_knn_index is not None
        template_ids, template_scores = tfidf_matching(
            query_features=query_features, object_repre=object_repre, top_n_templates=top_n_templates,
            visual_words_knn_index=visual_words_knn_index)
    else:
        raise ValueError(f"Unknown matching type '{matching_type}'.")
    logger.info(f"Matched templates: {list(misc.tensor_to_array(template_ids))}")
    return template_ids, template_scores
