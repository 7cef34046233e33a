"""Stage-wise parity checker of the batched CUDA pipeline against the oracle.  TEST INFRASTRUCTURE ONLY.

Used by tests/, `__graft_entry__.smoke()` and the pre-timing spot check of `bench.py` - never by the product.
Implements SURVEY.md §8c(iv): every stage is compared with the ORACLE run on that stage's inputs as the CUDA
path saw them (chained inputs), and the end-to-end agreement through the fp16 ViT is reported as a rate:

  stage 1  ViT + mask filter + sampling + PCA   (utils/dinov2_utils.py:115-158, utils/feature_util.py:75-131,
           utils/projector_util.py:66-88)        oracle on the same image/mask -> query points bit-exact,
                                                 descriptors within a relative Frobenius tolerance
  stage 2  tf-idf template retrieval            (utils/template_util.py:126-176) oracle on the CUDA path's own
           descriptors (fp16 rows, read back)    -> template ids bit-exact where the score gap allows
  stage 3  cyclic buddies per retrieved template (utils/corresp_util.py:34-70, 135-155), for the CUDA path's OWN
           template ids: (a) the two 1-NN searches vs the oracle on the same descriptors -> ids equal wherever the
           oracle's fp64 margin allows (asserted per query), distances within 1e-3; (b) cycle distances, top-k and
           gathers vs the oracle on the CUDA path's own 1-NN ids -> 2D ids, 3D ids, distances bit-exact (asserted)
  stage 4  full-bank k-NN (benchmark search K4)  (utils/knn_util.py:65-106 over all feat_vectors) for a few
           queries, bank streamed block-wise from HBM -> ids bit-exact outside the tie margin, d within 1e-3

The bank is read through the packed `ObjectIndex` (fp16 rows are lossless for fp16-representable banks), so the
check also works for banks that only ever existed on the device (the 10 k-template bank of BASELINE configs[2]).
"""

from __future__ import annotations

from typing import Any, Dict, Optional

import torch

from . import corresp as ocorresp
from . import feature as ofeature
from . import knn as oknn
from . import pca as opca
from . import template as otemplate
from . import vit as ovit

MARGIN = 1e-4          # relative k-NN gap above which index equality is required
DIST_RTOL = 1e-3       # north_star: descriptor distances within 1e-3 relative


def _index_cpu(index: Any) -> Dict[str, torch.Tensor]:
    """Small CPU copies of the index parts every check needs (NOT the bank rows)."""
    cache = getattr(index, "_oracle_cpu", None)
    if cache is None:
        cache = {
            "centroids": index.centroids16.float().cpu(),
            "idfs": index.idfs.cpu(),
            "template_descs": index.template_descs.cpu(),
            "tpl_off": index.tpl_off.cpu().to(torch.int64),
            "feat_perm": index.feat_perm.cpu() if index.feat_perm is not None else None,
        }
        index._oracle_cpu = cache
    return cache


def descriptor_stage(pipe: Any, crop: int, image: torch.Tensor, mask: torch.Tensor, state_dict: Dict, arch: Any,
                     layer: int, pca_dict: Optional[Dict], facet: str = "token", apply_norm: bool = True,
                     desc: Optional[torch.Tensor] = None) -> Dict:
    """Stage 1 for one crop.  image [3,H,W] fp32 CPU, mask [H,W] bool CPU.  `desc`: the [B*stride, d] tensor the
    pipeline wrote its projected descriptors to (default: its own fp32 / fp16 buffer)."""
    size = (pipe.crop_w, pipe.crop_h)
    fmap = ovit.extract(state_dict, arch, image.unsqueeze(0), layer=layer, facet=facet,
                        apply_norm=apply_norm)["feature_maps"][0]
    grid = ofeature.generate_grid_points(size, float(size[0] // pipe.wp))
    qp = ofeature.filter_points_by_mask(grid, mask)
    n = int(pipe.q_count[crop].item())
    points_equal = n == qp.shape[0] and torch.equal(pipe.q_points[crop, :n].cpu(), qp)
    res = {"n_queries": n, "points_equal": bool(points_equal), "oracle_points": qp}
    if n == 0 or not points_equal:
        res["rel_err"] = float("nan")
        return res
    feats = ofeature.sample_feature_map_at_points(fmap, qp, size).contiguous()
    if pca_dict is not None:
        feats = opca.project_features(feats, [pca_dict]).contiguous()
    s = pipe.stride
    src = desc if desc is not None else (pipe.proj32 if pipe.proj32 is not None else pipe.proj16)
    ours = src[crop * s: crop * s + n, : feats.shape[1]].float().cpu()
    res["rel_err"] = float(torch.linalg.norm(ours - feats) / torch.linalg.norm(feats))
    cos = torch.nn.functional.cosine_similarity(ours, feats, dim=1)
    res["min_cos"] = float(cos.min())
    res["oracle_desc"] = feats
    return res


def retrieval_stage(index: Any, engine: Any, crop: int, q_points: torch.Tensor, q_desc: torch.Tensor, top_k: int
                    ) -> Dict:
    """Stages 2 and 3 for one crop of a RetrievalEngine run, oracle fed with the descriptors the CUDA path matched.

    q_points [n,2] fp32 CPU, q_desc [n, d_padded] fp32 CPU (the fp16 rows the kernels read, widened).  Chained:
      2   tf-idf + cosine scores + top-N     vs otemplate.tfidf_matching on q_desc
      3a  the two 1-NN searches per pair     vs oknn.knn_l2 on (q_desc, the template's rows): ids equal wherever the
                                             oracle's fp64 margin > MARGIN (asserted), distances within DIST_RTOL
      3b  cycle distance, top-k, gathers     vs ocorresp.cyclic_buddies_from_ids on the CUDA path's OWN 1-NN ids:
                                             2D ids, 3D ids and distances bit-exact (asserted)
    """
    ix = _index_cpu(index)
    out = engine.out
    n = q_points.shape[0]
    topn = out.template_ids.shape[1]
    res: Dict[str, Any] = {"n_queries": n}
    if n == 0:
        assert bool((out.count[crop] == 0).all().item())
        res.update(templates_equal=True, templates_sure=True, pairs=0, pairs_exact=0, nn_checked=0, nn_sure_frac=1.0,
                   nn_equal_frac=1.0, corr_agree=1.0)
        return res
    ids, scores, tfidf, cos = otemplate.tfidf_matching(
        q_desc, ix["centroids"], ix["idfs"], ix["template_descs"], topn, index.tfidf_knn_k, index.tfidf_knn_metric,
        index.tfidf_soft_assign, index.tfidf_soft_sigma_squared)
    ours_ids = out.template_ids[crop].cpu()
    ours_scores = out.template_scores[crop].cpu()
    srt = torch.sort(cos, descending=True).values[: topn + 1]
    gap = float((srt[:-1] - srt[1:]).min()) if srt.numel() > 1 else float("inf")
    res["template_gap"] = gap
    res["templates_equal"] = bool(torch.equal(ours_ids, ids))
    res["templates_sure"] = gap > 1e-5
    res["template_score_err"] = float((ours_scores - cos[ours_ids]).abs().max())
    res["tfidf_err"] = float((out.query_tfidf[crop].cpu() - tfidf).abs().max())
    pairs = exact = 0
    nn_total = nn_sure = nn_equal = 0
    agree = total = 0
    nn_dist_err = 0.0
    for j in range(topn):
        t = int(ours_ids[j])
        r0, r1 = int(ix["tpl_off"][t]), int(ix["tpl_off"][t + 1])
        cnt = int(out.count[crop, j].item())
        if r1 == r0:
            assert cnt == 0, "a template without bank rows must yield no correspondences"
            continue
        p = crop * topn + j
        rows = index.bank16[r0:r1].float().cpu()
        g_q2o = engine.q2o_i[p * engine.stride: p * engine.stride + n, 0].cpu()
        g_o2q = engine.o2q_i[p * engine.max_p: p * engine.max_p + (r1 - r0), 0].cpu()
        g_q2o_d = engine.q2o_d[p * engine.stride: p * engine.stride + n, 0].cpu()
        # 3a: the two 1-NN searches
        for ours_i, (qq, xx), ours_d in ((g_q2o, (q_desc, rows), g_q2o_d), (g_o2q, (rows, q_desc), None)):
            rd, ri = oknn.knn_l2(qq, xx, 1)
            sure_m = oknn.topk_margin(qq, xx, 1)[:, 0] > MARGIN
            same = ours_i == ri[:, 0]
            assert bool(same[sure_m].all()), f"crop {crop} template {t}: 1-NN ids differ outside the tie margin"
            nn_total += same.numel(); nn_sure += int(sure_m.sum()); nn_equal += int(same.sum())
            if ours_d is not None:
                err = ((ours_d - rd[:, 0]).abs() / rd[:, 0].clamp_min(1e-6)).max()
                nn_dist_err = max(nn_dist_err, float(err))
        assert nn_dist_err <= DIST_RTOL, f"1-NN distances off by {nn_dist_err:.2e} relative"
        # 3b: cyclic buddies from the CUDA path's own ids
        q_ids, o_ids, dists, conf = ocorresp.cyclic_buddies_from_ids(q_points, g_q2o, g_o2q, top_k)
        feat_ids = o_ids + r0
        if ix["feat_perm"] is not None:
            feat_ids = ix["feat_perm"][feat_ids]
        g_q = out.query_ids[crop, j, :cnt].cpu()
        g_v = out.vertex_ids[crop, j, :cnt].cpu()
        g_d = out.dists[crop, j, :cnt].cpu()
        same = cnt == q_ids.shape[0] and torch.equal(g_q, q_ids) and torch.equal(g_v, feat_ids) \
            and torch.equal(g_d, dists)
        assert same, f"crop {crop} template {t}: cyclic-buddies outputs differ from the oracle on the same 1-NN ids"
        assert torch.allclose(out.scores[crop, j, :cnt].cpu(), conf, equal_nan=True)
        assert torch.equal(out.coord_2d[crop, j, :cnt].cpu(), q_points[q_ids])
        pairs += 1
        exact += int(same)
        # full-oracle comparison of the pair (1-NN + cycle by the oracle), as a rate
        o_q, o_o, _, _ = ocorresp.cyclic_buddies_matching(q_points, q_desc, rows, top_k)
        o_f = o_o + r0
        if ix["feat_perm"] is not None:
            o_f = ix["feat_perm"][o_f]
        if cnt == o_q.shape[0]:
            agree += int(((g_q == o_q) & (g_v == o_f)).sum())
        total += o_q.shape[0]
    res.update(pairs=pairs, pairs_exact=exact, nn_checked=nn_total, nn_sure_frac=nn_sure / max(nn_total, 1),
               nn_equal_frac=nn_equal / max(nn_total, 1), nn_dist_rel_err=nn_dist_err,
               corr_agree=agree / max(total, 1))
    return res


def bank_blocks(bank16: torch.Tensor, block_rows: int = 1 << 19):
    """(first_row, fp32 CPU rows) blocks of a device-resident fp16 bank."""
    for s in range(0, bank16.shape[0], block_rows):
        yield s, bank16[s:s + block_rows].float().cpu()


def full_bank_knn_stage(bank16: torch.Tensor, q_desc: torch.Tensor, ours_d: torch.Tensor, ours_i: torch.Tensor,
                        k: int) -> Dict:
    """Stage 4: oracle full-bank search for the given queries (fp32 CPU rows) vs the CUDA results."""
    rd, ri = oknn.knn_l2_blocked(q_desc, bank_blocks(bank16), k + 1)
    gap = (rd[:, 1:] - rd[:, :-1]) / rd[:, :-1].clamp_min(1e-12)          # gaps d_{j+1} - d_j, j = 0..k-1
    prev = torch.cat([torch.full_like(gap[:, :1], float("inf")), gap[:, :-1]], dim=1)
    sure = torch.minimum(gap, prev) > MARGIN
    rd, ri = rd[:, :k], ri[:, :k]
    same = ours_i.cpu() == ri
    assert bool(same[sure].all()), "full-bank k-NN ids differ outside the tie margin"
    ok = torch.isfinite(rd)
    derr = ((ours_d.cpu() - rd).abs()[ok] / rd[ok].clamp_min(1e-6)).max() if ok.any() else torch.tensor(0.0)
    assert float(derr) <= DIST_RTOL, f"full-bank k-NN distances off by {float(derr):.2e} relative"
    return {"queries": int(q_desc.shape[0]), "ids_equal": float(same.float().mean()),
            "sure_frac": float(sure.float().mean()), "dist_rel_err": float(derr)}


def end_to_end_agreement(index: Any, out: Any, crop: int, q_points: torch.Tensor, oracle_desc: torch.Tensor,
                         top_k: int) -> Dict:
    """End-to-end rate (§8c(iv)): oracle descriptors (fp32 ViT) through the oracle matcher vs the CUDA outputs."""
    ix = _index_cpu(index)
    topn = out.template_ids.shape[1]
    d_pad = index.dim_padded
    q = oracle_desc
    if q.shape[1] < d_pad:
        q = torch.nn.functional.pad(q, (0, d_pad - q.shape[1]))
    ids, _, _, _ = otemplate.tfidf_matching(
        q, ix["centroids"], ix["idfs"], ix["template_descs"], topn, index.tfidf_knn_k, index.tfidf_knn_metric,
        index.tfidf_soft_assign, index.tfidf_soft_sigma_squared)
    ours_ids = out.template_ids[crop].cpu()
    tpl_agree = float((ours_ids == ids).float().mean())
    agree = total = 0
    for j in range(topn):
        if int(ours_ids[j]) != int(ids[j]):
            continue
        t = int(ids[j])
        r0, r1 = int(ix["tpl_off"][t]), int(ix["tpl_off"][t + 1])
        if r1 == r0:
            continue
        rows = index.bank16[r0:r1].float().cpu()
        q_ids, o_ids, _, _ = ocorresp.cyclic_buddies_matching(q_points, q, rows, top_k)
        cnt = int(out.count[crop, j].item())
        if cnt != q_ids.shape[0]:
            continue
        feat_ids = o_ids + r0
        if ix["feat_perm"] is not None:
            feat_ids = ix["feat_perm"][feat_ids]
        # order-insensitive: fraction of the oracle's (2D id, 3D id) pairs the CUDA path also produced
        ours_pairs = set(zip(out.query_ids[crop, j, :cnt].cpu().tolist(), out.vertex_ids[crop, j, :cnt].cpu().tolist()))
        agree += sum((a, b) in ours_pairs for a, b in zip(q_ids.tolist(), feat_ids.tolist()))
        total += cnt
    return {"template_id_agreement": tpl_agree, "corresp_pair_agreement": agree / total if total else float("nan")}
