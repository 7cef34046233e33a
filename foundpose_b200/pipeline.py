"""Batched, synchronisation-free execution of the per-crop hot path on one GPU.

The reference processes one crop at a time with 11 host round-trips per crop
(scripts/infer.py:467-545).  Here a batch of crops goes through

    ViT -> mask filter -> feature sampling -> PCA -> visual-word k-NN -> tf-idf -> cosine scores
        -> top-N templates -> per-(crop, template) 1-NN both ways -> cyclic buddies -> gathers

as a fixed sequence of kernel launches on one stream with every intermediate resident in HBM and no
device->host copy until the caller reads the results.

`ObjectIndex` is the packed HBM layout of one object's representation: fp16 bank rows sorted by
template with CSR offsets (instead of the reference's T separate faiss indexes and
`feat_to_template_ids == t` scans, scripts/infer.py:224-239, utils/corresp_util.py:110-113),
precomputed ||x||^2, fp16 visual words, idfs, template descriptors and their norms, vertices.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Tuple

import torch

from foundpose_b200 import _native


# Descriptor rows are padded to a multiple of 128 columns everywhere: the PCA GEMM writes N % 128 == 0 columns
# (PCAProjector.device_state) and the bank rows it is matched against must have the same width.
DESC_PAD = 128
# Bag-of-words scores on the tensor cores: both sides are scaled by 2^10 before the hi/lo split (keeps the lo parts
# of unit-vector entries out of the fp16 subnormals); inner products come back multiplied by 2^20.
BOW_SCALE = 1024.0
BOW_CHUNK_ROWS = 256      # templates per k-NN item (one bank tile): T / 256 SMs stream the descriptors in parallel


def _pad_cols(x: torch.Tensor, mult: int = DESC_PAD) -> torch.Tensor:
    d = x.shape[1]
    dp = (d + mult - 1) // mult * mult
    if dp == d:
        return x.contiguous()
    return torch.nn.functional.pad(x, (0, dp - d)).contiguous()


class ObjectIndex:
    """Packed device-resident index of a FeatureBasedObjectRepre (fields: utils/repre_util.py:34-83)."""

    def __init__(self, repre: Any, device: torch.device, num_templates: Optional[int] = None) -> None:
        dev = torch.device(device)
        self.device = dev
        # fp16 rows (a bank that was packed before, e.g. received through broadcast_object_repre) are taken as they
        # are; fp32 rows (repre.pth, utils/repre_util.py:143-210) are converted once.
        packed16 = repre.feat_vectors.dtype == torch.float16
        feat = repre.feat_vectors.to(dev) if packed16 else repre.feat_vectors.to(dev, torch.float32)
        tpl_ids = repre.feat_to_template_ids.to(dev).to(torch.int64)
        self.feat_dim = feat.shape[1]
        # Bank rows must be contiguous per template. gen_repre builds them that way
        # (scripts/gen_repre.py:187-190, 214); otherwise sort once and remember the permutation.
        if tpl_ids.numel() > 1 and bool((tpl_ids[1:] < tpl_ids[:-1]).any()):
            perm = torch.sort(tpl_ids, stable=True).indices
            self.feat_perm: Optional[torch.Tensor] = perm.contiguous()
            feat = feat[perm]
            tpl_ids = tpl_ids[perm]
        else:
            self.feat_perm = None
        if num_templates is None:
            if getattr(repre, "template_descs", None) is not None:
                num_templates = int(repre.template_descs.shape[0])
            else:
                num_templates = int(tpl_ids.max().item()) + 1 if tpl_ids.numel() else 0
        self.num_templates = num_templates
        counts = torch.bincount(tpl_ids, minlength=num_templates)[:num_templates]
        off = torch.zeros(num_templates + 1, dtype=torch.int64, device=dev)
        off[1:] = torch.cumsum(counts, 0)
        self.tpl_off = off.to(torch.int32).contiguous()
        self.max_template_rows = int(counts.max().item()) if num_templates else 0
        self.bank16 = _pad_cols(feat) if packed16 else _native.convert_rows_f16(_pad_cols(feat))
        self.bank_sqnorm = _native.row_sqnorm_f16(self.bank16)
        self.dim_padded = self.bank16.shape[1]
        self.vertices = repre.vertices.to(dev, torch.float32).contiguous() if repre.vertices is not None else None

        self.centroids16 = self.centroid_sqnorm = self.idfs = self.template_descs = self.desc_norm = None
        if getattr(repre, "feat_cluster_centroids", None) is not None:
            cent = repre.feat_cluster_centroids.to(dev, torch.float32)
            opts_ = getattr(repre, "template_desc_opts", None)
            cosine_words = opts_ is not None and opts_.tfidf_knn_metric == "cosine"
            # metric "cosine": visual words are L2-normalised at fit time (knn_util.py:61-64).
            self.centroids16 = _native.convert_rows_f16(_pad_cols(cent), l2_normalize=cosine_words)
            self.centroid_sqnorm = _native.row_sqnorm_f16(self.centroids16)
        if getattr(repre, "feat_cluster_idfs", None) is not None:
            self.idfs = repre.feat_cluster_idfs.to(dev, torch.float32).contiguous()
        self.desc_split16: Optional[torch.Tensor] = None     # built on first use by a RetrievalEngine
        self.idfs_finite = True
        if self.idfs is not None:
            self.idfs_finite = bool(torch.isfinite(self.idfs).all().item())
        if getattr(repre, "template_descs", None) is not None:
            self.template_descs = repre.template_descs.to(dev, torch.float32).contiguous()
            self.desc_norm = _native.row_norm_f32(self.template_descs)
        opts = getattr(repre, "template_desc_opts", None)
        self.tfidf_knn_k = opts.tfidf_knn_k if opts is not None else 3
        self.tfidf_knn_metric = opts.tfidf_knn_metric if opts is not None else "l2"
        if self.tfidf_knn_metric not in ("l2", "cosine"):
            raise ValueError(f"Metric {self.tfidf_knn_metric} is not supported.")
        self.tfidf_soft_assign = bool(opts.tfidf_soft_assign) if opts is not None else False
        self.tfidf_soft_sigma_squared = float(opts.tfidf_soft_sigma_squared) if opts is not None else 10.0

    @property
    def num_words(self) -> int:
        return int(self.centroids16.shape[0])

    def split_descs(self) -> torch.Tensor:
        """Unit-norm template descriptors as fp16 [T, 3W] = [hi | lo | hi] * 2^10 (fp_split_rows_f16): the index of
        the tensor-core bag-of-words scoring."""
        if self.desc_split16 is None:
            self.desc_split16 = _native.split_rows_f16(self.template_descs, 0, True, BOW_SCALE)
        return self.desc_split16


def get_object_index(repre: Any, device: torch.device) -> ObjectIndex:
    """ObjectIndex cached on the repre object (built once per object, like infer.py:215-239)."""
    cache = getattr(repre, "_b200_index", None)
    if cache is None or cache.device != torch.device(device):
        cache = ObjectIndex(repre, device)
        try:
            object.__setattr__(repre, "_b200_index", cache)
        except Exception:
            pass
    return cache


@dataclass
class MatchOutputs:
    """Device tensors produced for a batch of B crops x N retrieved templates."""

    template_ids: torch.Tensor      # [B, N] int64
    template_scores: torch.Tensor   # [B, N] fp32
    count: torch.Tensor             # [B, N] int32: valid correspondences per pair
    query_ids: torch.Tensor         # [B, N, K] int64  (coord_2d_ids)
    vertex_ids: torch.Tensor        # [B, N, K] int64  (nn_vertex_ids)
    dists: torch.Tensor             # [B, N, K] fp32   (cyclic distances)
    scores: torch.Tensor            # [B, N, K] fp32   (coord_conf)
    coord_2d: torch.Tensor          # [B, N, K, 2]
    coord_3d: torch.Tensor          # [B, N, K, 3]
    query_tfidf: torch.Tensor       # [B, W]
    cos_sims: torch.Tensor          # [B, T]
    word_ids: torch.Tensor          # [B*stride, k]
    word_dists: torch.Tensor        # [B*stride, k] squared L2 to the visual words


class RetrievalEngine:
    """Template retrieval + cyclic-buddies matching for B crops with a fixed per-crop row stride."""

    def __init__(self, index: ObjectIndex, batch: int, stride: int, top_n_templates: int, top_k_buddies: int,
                 knn_k: Optional[int] = None) -> None:
        assert index.centroids16 is not None and index.idfs is not None and index.template_descs is not None, \
            "Template descriptors need to be tfidf."
        self.index = index
        dev = index.device
        self.batch, self.stride = batch, stride
        self.topn = min(top_n_templates, index.num_templates)
        self.top_k = top_k_buddies
        self.knn_k = knn_k if knn_k is not None else index.tfidf_knn_k
        rows = batch * stride
        self.rows = rows
        self.max_p = max(index.max_template_rows, 1)
        npairs = batch * self.topn
        f32, i64, i32 = torch.float32, torch.int64, torch.int32
        self.q_start = (torch.arange(batch, dtype=i32, device=dev) * stride).contiguous()
        self.q_sqnorm = torch.empty(rows, dtype=f32, device=dev)
        self.items_words = _native.new_knn_items(_native.knn_num_items(rows), dev)
        self.n_items_words = _native.knn_num_items(rows)
        _native.knn_items_dense(self.items_words, rows, 0, index.num_words)
        self.word_d = torch.empty((rows, self.knn_k), dtype=f32, device=dev)
        self.word_i = torch.empty((rows, self.knn_k), dtype=i64, device=dev)
        self.tfidf = torch.empty((batch, index.num_words), dtype=f32, device=dev)
        self.cos = torch.empty((batch, index.num_templates), dtype=f32, device=dev)
        self.top_scores = torch.empty((batch, self.topn), dtype=f32, device=dev)
        self.top_ids = torch.empty((batch, self.topn), dtype=i64, device=dev)
        # Template scoring (utils/template_util.py:160-176) as ONE inner-product top-N search on the tensor cores:
        # cos(q, d) = <q/||q||, d/||d||> with both unit vectors split into fp16 hi + lo parts (fp32-accurate), the
        # k-NN kernel's epilogue keeps the N best templates per crop, so the [B, T] score matrix never exists.
        # Kept on the fp32 CUDA-core path: N > 16, a vocabulary whose 3W is not a multiple of 64, non-finite idfs (the reference's NaN scores) and
        # callers that want the full score matrix (`full_scores`).
        self.full_scores = False
        self.bow_tensor = self.topn <= 16 and index.idfs_finite and (3 * index.num_words) % 64 == 0
        if self.bow_tensor:
            T, W = index.num_templates, index.num_words
            self.bow_chunks = (T + BOW_CHUNK_ROWS - 1) // BOW_CHUNK_ROWS
            self.bow_qblocks = _native.knn_num_items(batch)
            self.bow_items = _native.new_knn_items(self.bow_qblocks * self.bow_chunks, dev)
            _native.knn_items_split(self.bow_items, batch, 0, T, self.bow_chunks, BOW_CHUNK_ROWS)
            self.bow_q16 = torch.empty((batch, 3 * W), dtype=torch.float16, device=dev)
            q_pad = self.bow_qblocks * 128
            self.bow_part_d = torch.empty((self.bow_chunks * q_pad, self.topn), dtype=f32, device=dev)
            self.bow_part_i = torch.full((self.bow_chunks * q_pad, self.topn), -1, dtype=i64, device=dev)
            index.split_descs()
        self.per_q = (stride + 127) // 128
        self.per_p = (self.max_p + 127) // 128
        self.items_q2o = _native.new_knn_items(npairs * self.per_q, dev)
        self.items_o2q = _native.new_knn_items(npairs * self.per_p, dev)
        self.q2o_d = torch.empty((npairs * stride, 1), dtype=f32, device=dev)
        self.q2o_i = torch.zeros((npairs * stride, 1), dtype=i64, device=dev)
        self.o2q_d = torch.empty((npairs * self.max_p, 1), dtype=f32, device=dev)
        self.o2q_i = torch.zeros((npairs * self.max_p, 1), dtype=i64, device=dev)
        k = top_k_buddies
        self.cyc_workspace = _native.cyclic_buddies_workspace(npairs, stride, top_k_buddies, dev)
        self.q_unit16 = self.q_unit_sqnorm = None
        self.out = MatchOutputs(
            template_ids=self.top_ids, template_scores=self.top_scores,
            count=torch.zeros((batch, self.topn), dtype=i32, device=dev),
            query_ids=torch.zeros((batch, self.topn, k), dtype=i64, device=dev),
            vertex_ids=torch.zeros((batch, self.topn, k), dtype=i64, device=dev),
            dists=torch.zeros((batch, self.topn, k), dtype=f32, device=dev),
            scores=torch.zeros((batch, self.topn, k), dtype=f32, device=dev),
            coord_2d=torch.zeros((batch, self.topn, k, 2), dtype=f32, device=dev),
            coord_3d=torch.zeros((batch, self.topn, k, 3), dtype=f32, device=dev),
            query_tfidf=self.tfidf, cos_sims=self.cos, word_ids=self.word_i, word_dists=self.word_d)

    def score_templates(self) -> None:
        """self.tfidf [B, W] -> top-N template ids / cosine scores (utils/template_util.py:160-176)."""
        ix = self.index
        if self.bow_tensor:
            _native.split_rows_f16(self.tfidf, 1, True, BOW_SCALE, out=self.bow_q16)
            _native.knn_search_items(self.bow_q16, None, ix.split_descs(), None, self.bow_items,
                                     self.bow_qblocks * self.bow_chunks, 1, self.topn, self.bow_part_d,
                                     self.bow_part_i)
            _native.knn_merge(self.bow_part_d, self.bow_part_i, self.bow_chunks, self.bow_qblocks * 128, self.batch,
                              self.topn, BOW_CHUNK_ROWS, ix.num_templates, True, self.top_scores, self.top_ids)
            self.top_scores.mul_(1.0 / (BOW_SCALE * BOW_SCALE))
        if self.full_scores or not self.bow_tensor:
            _native.bow_scores(ix.template_descs, ix.desc_norm, self.tfidf, self.cos)
            if not self.bow_tensor:
                _native.topk_rows(self.cos, self.topn, self.top_scores, self.top_ids)

    def match(self, feat16: torch.Tensor, points: torch.Tensor, q_count: torch.Tensor) -> MatchOutputs:
        """feat16 [B*stride, dpad] f16 query descriptors, points [B, stride, 2], q_count int32 [B]."""
        ix = self.index
        assert feat16.shape == (self.rows, ix.dim_padded), (feat16.shape, self.rows, ix.dim_padded)
        _native.row_sqnorm_f16(feat16, self.q_sqnorm)
        # K1: nearest visual words of every query (template_util.py:13-29).
        if ix.tfidf_knn_metric == "l2":
            _native.knn_search_items(feat16, self.q_sqnorm, ix.centroids16, ix.centroid_sqnorm, self.items_words,
                                     self.n_items_words, 0, self.knn_k, self.word_d, self.word_i)
        else:
            # cosine: normalised copies of the queries (one fused pass over the projected descriptors), inner-product
            # search, distance = 1 - similarity (knn_util.py:84-98).
            if self.q_unit16 is None:
                self.q_unit16 = torch.empty_like(feat16)
                self.q_unit_sqnorm = torch.empty_like(self.q_sqnorm)
            _native.unit_rows_f16(feat16, self.q_unit16, self.q_unit_sqnorm)
            _native.knn_search_items(self.q_unit16, self.q_unit_sqnorm, ix.centroids16, ix.centroid_sqnorm,
                                     self.items_words, self.n_items_words, 1, self.knn_k, self.word_d, self.word_i)
            torch.sub(1.0, self.word_d, out=self.word_d)
        _native.calc_tfidf(self.word_i, self.word_d, self.q_start, q_count, ix.idfs, ix.tfidf_soft_assign,
                           ix.tfidf_soft_sigma_squared, True, self.tfidf)
        self.score_templates()
        # K2 / K3: 1-NN in both directions for every (crop, retrieved template) pair.
        _native.build_pair_items(self.top_ids, self.topn, ix.tpl_off, self.q_start, q_count, self.stride,
                                 self.max_p, self.items_q2o, self.items_o2q)
        npairs = self.batch * self.topn
        _native.knn_search_items(feat16, self.q_sqnorm, ix.bank16, ix.bank_sqnorm, self.items_q2o,
                                 npairs * self.per_q, 0, 1, self.q2o_d, self.q2o_i)
        _native.knn_search_items(ix.bank16, ix.bank_sqnorm, feat16, self.q_sqnorm, self.items_o2q,
                                 npairs * self.per_p, 0, 1, self.o2q_d, self.o2q_i)
        o = self.out
        _native.cyclic_buddies(points, self.q_start, q_count, self.q2o_i, self.o2q_i, self.top_ids, self.topn,
                               ix.tpl_off, ix.feat_perm, ix.vertices, self.stride, self.max_p, self.top_k,
                               o.query_ids, o.vertex_ids, o.dists, o.scores, o.coord_2d, o.coord_3d, o.count,
                               self.cyc_workspace)
        return o


class CropBatchPipeline:
    """Full hot path for a batch of crops: images + masks in, correspondences out."""

    def __init__(self, extractor: Any, index: ObjectIndex, projectors: Optional[List[Any]], batch: int,
                 crop_size: Tuple[int, int] = (420, 420), grid_cell_size: float = 14.0,
                 top_n_templates: int = 5, top_k_buddies: int = 300, keep_f32_descriptors: bool = False) -> None:
        """keep_f32_descriptors: also write the projected descriptors in fp32 (`proj32`, 88 MB per 64 crops at
        d = 384); the matching stages only read the fp16 copy."""
        from foundpose_b200.utils import feature_util

        self.extractor = extractor
        self.index = index
        dev = index.device
        self.batch = batch
        self.crop_w, self.crop_h = crop_size
        ps = extractor.patch_size
        self.hp, self.wp = self.crop_h // ps, self.crop_w // ps
        self.grid_points = feature_util.generate_grid_points(crop_size, grid_cell_size).to(dev).contiguous()
        self.stride = self.grid_points.shape[0]
        d_vit = extractor.arch.embed_dim
        self.projector = None
        if projectors:
            from foundpose_b200.utils import projector_util

            # a chain of projectors (project_features applies them in turn) collapses to one affine map: one GEMM
            self.projector = projector_util.compose_projectors(list(projectors)).device_state(dev)
        rows = batch * self.stride
        f32, f16, i32 = torch.float32, torch.float16, torch.int32
        self.tokens = torch.empty((batch, self.hp * self.wp, d_vit), dtype=f32, device=dev)
        self.q_points = torch.zeros((batch, self.stride, 2), dtype=f32, device=dev)
        self.q_ids = torch.zeros((batch, self.stride), dtype=i32, device=dev)
        self.q_count = torch.zeros((batch,), dtype=i32, device=dev)
        self.sampled16 = torch.zeros((rows, d_vit), dtype=f16, device=dev)
        if self.projector is not None:
            d_out = self.projector["components16"].shape[0]
            self.proj32 = torch.empty((rows, d_out), dtype=f32, device=dev) if keep_f32_descriptors else None
            self.proj16 = torch.empty((rows, d_out), dtype=f16, device=dev)
            assert d_out == index.dim_padded, (d_out, index.dim_padded)
        else:
            assert d_vit == index.dim_padded, (d_vit, index.dim_padded)
            self.proj32 = None
            self.proj16 = self.sampled16
        self.engine = RetrievalEngine(index, batch, self.stride, top_n_templates, top_k_buddies)

    @torch.no_grad()
    def run(self, images: torch.Tensor, masks_u8: torch.Tensor, desc_out: Optional[torch.Tensor] = None
            ) -> MatchOutputs:
        """images fp32 [B,3,H,W] in [0,1], masks uint8 [B,H,W]; both on the pipeline's device.

        desc_out (optional, fp16 [B*stride, dim_padded]): where the projected query descriptors are written (and
        matched from) instead of the pipeline's own buffer - e.g. a slice of a larger query set that is searched
        against the whole bank afterwards."""
        assert images.shape[0] == self.batch and masks_u8.shape[0] == self.batch
        desc = self.proj16 if desc_out is None else desc_out
        assert desc.shape == self.proj16.shape and desc.dtype == torch.float16 and desc.is_contiguous()
        self.extractor.forward_tokens(images, want_f16=False, want_cls=False, out_tokens=self.tokens)
        _native.filter_points_by_mask(self.grid_points, masks_u8, self.q_points, self.q_ids, self.q_count)
        _native.sample_features(self.tokens, self.hp, self.wp, self.q_points, self.q_count,
                                float(self.crop_w), float(self.crop_h), None, self.sampled16)
        if self.projector is not None:
            _native.pca_project(self.sampled16, self.projector["components16"], self.projector["bias"],
                                self.proj32, desc)
        elif desc_out is not None:
            desc.copy_(self.sampled16)
        return self.engine.match(desc, self.q_points, self.q_count)


def outputs_to_corresp_list(out: MatchOutputs, crop: int, debug: bool = False) -> List[Dict[str, torch.Tensor]]:
    """The reference's list-of-dicts result for one crop (utils/corresp_util.py:147-165)."""
    res = []
    counts = out.count[crop].tolist()
    for j, n in enumerate(counts):
        item = {
            "template_id": out.template_ids[crop, j],
            "template_score": out.template_scores[crop, j],
            "coord_2d": out.coord_2d[crop, j, :n],
            "coord_2d_ids": out.query_ids[crop, j, :n],
            "coord_3d": out.coord_3d[crop, j, :n],
            "coord_conf": out.scores[crop, j, :n],
            "nn_vertex_ids": out.vertex_ids[crop, j, :n],
        }
        if debug:
            item.update({"nn_dists": out.dists[crop, j, :n], "nn_indices": out.vertex_ids[crop, j, :n]})
        res.append(item)
    return res
