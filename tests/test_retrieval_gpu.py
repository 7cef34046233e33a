"""GPU parity of the sampling / PCA / k-NN / tf-idf / cyclic-buddies stages (A5-A16) against the
CPU oracle and the golden outputs of the reference's own modules.

Bars: bit-exact for indices (k-NN ids where the oracle margin > 1e-4 relative, template ids,
correspondence ids with canonical tie order); distances within 1e-3 relative (north_star);
fp32 streaming stages within the tolerance written in each test.
"""
import os

import pytest
import torch

from foundpose_b200 import synthetic

pytestmark = pytest.mark.gpu
GOLD_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.pt")


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD_PATH, weights_only=False)


def _small_repre(gold, device="cpu"):
    from foundpose_b200.utils import repre_util

    bank = synthetic.make_bank_tensors(num_templates=24, patches_per_template=48, feat_dim=64, num_words=32,
                                       seed=gold["bank/seed"], ragged=True)
    return repre_util.FeatureBasedObjectRepre(
        vertices=bank["vertices"], feat_vectors=bank["feat_vectors"],
        feat_to_template_ids=bank["feat_to_template_ids"], feat_to_vertex_ids=bank["feat_to_vertex_ids"],
        feat_cluster_centroids=bank["feat_cluster_centroids"], feat_cluster_idfs=gold["bank/idfs"],
        template_descs=gold["bank/template_descs"], template_desc_opts=repre_util.TemplateDescOpts(),
        feat_opts=repre_util.FeatureOpts("dinov2_vits14-reg")), bank


def test_grid_filter_sample_vs_golden(gold):
    from foundpose_b200.utils import feature_util

    grid14 = feature_util.generate_grid_points((420, 420), 14.0)
    assert torch.equal(grid14, gold["feature/grid14"])
    mask = synthetic.make_masks(1, (420, 420), seed=gold["feature/mask_seed"])[0]
    qp = feature_util.filter_points_by_mask(grid14.cuda(), mask.cuda())
    assert torch.equal(qp.cpu(), gold["feature/filtered14"])          # bit-exact selection + order
    empty = feature_util.filter_points_by_mask(grid14.cuda(), torch.zeros(420, 420, dtype=torch.bool).cuda())
    assert empty.shape == (0, 2)
    fmap = torch.randn(48, 30, 30, generator=torch.Generator().manual_seed(gold["feature/fmap_seed"]))
    s = feature_util.sample_feature_map_at_points(fmap.cuda(), qp, (420, 420))
    # bilinear weights in fp32: |err| <= 1e-5 (absolute, values are O(1))
    assert (s.cpu() - gold["feature/sampled14"]).abs().max().item() <= 1e-5
    s2 = feature_util.sample_feature_map_at_points(fmap.cuda(), gold["feature/random_points"].cuda(), (420, 420))
    assert (s2.cpu() - gold["feature/sampled_random"]).abs().max().item() <= 1e-5


def test_pca_vs_golden(gold):
    from foundpose_b200.utils import projector_util

    pdict = synthetic.make_pca(128, 64, seed=41)
    proj = projector_util.projector_from_tensordict(pdict)
    x = synthetic.fp16_representable(torch.randn(300, 128, generator=torch.Generator().manual_seed(gold["pca/in_seed"])))
    out = projector_util.project_features(x.cuda(), [proj])
    assert out.shape == (300, 64)
    # fp16-representable inputs: only the accumulation order differs -> 1e-3 relative to max|ref|.
    ref = gold["pca/out"]
    assert (out.cpu() - ref).abs().max().item() <= 1e-3 * ref.abs().max().item()
    back = projector_util.projector_to_tensordict(proj)
    assert torch.equal(back["pca_projector"]["components"], pdict["pca_projector"]["components"])


@pytest.mark.parametrize("nq,nb,dim,k", [(150, 32, 64, 3), (1000, 2048, 256, 3), (900, 1024, 384, 1),
                                         (37, 5000, 128, 5), (129, 129, 64, 8), (5, 3, 64, 1),
                                         (100, 40000, 128, 5), (300, 70001, 64, 1),    # split-bank path
                                         # pair kernel (cta_group::2, >= 74 query blocks x >= 8192 bank rows):
                                         (9500, 9000, 384, 5),      # queries resident in smem, one partial wave
                                         (21541, 9000, 64, 5),      # one whole wave + a split tail wave + merge
                                         (9600, 8200, 640, 3),      # d > 576: queries streamed with the bank
                                         (9473, 8193, 128, 1)])     # ragged last item / last tile, k = 1
def test_knn_l2_vs_oracle(nq, nb, dim, k):
    from foundpose_b200.utils import knn_util
    from oracle import knn as oknn

    g = torch.Generator().manual_seed(nq + nb)
    bank = synthetic.fp16_representable(torch.randn(nb, dim, generator=g))
    q = synthetic.make_query_features(nq, dim, bank, seed=5)
    index = knn_util.KNN(k=k, metric="l2")
    index.fit(bank.cuda())
    d, i = index.search(q.cuda())
    assert d.dtype == torch.float32 and i.dtype == torch.int64 and d.is_cuda
    rd, ri = oknn.knn_l2(q, bank, k)
    margin = oknn.topk_margin(q, bank, k)
    sure = margin > 1e-4
    assert torch.equal(i.cpu()[sure], ri[sure]), "index mismatch outside the tie margin"
    assert sure.float().mean().item() > 0.95
    ok = torch.isfinite(rd)
    assert ((d.cpu() - rd).abs()[ok] <= 1e-3 * rd.abs()[ok] + 1e-4).all()   # 1e-3 relative (north_star)
    assert bool((d[:, 1:] >= d[:, :-1]).all())
    # CPU tensors are accepted like the reference's KNN (results come back on the CPU).
    d2, i2 = index.search(q)
    assert not d2.is_cuda and torch.equal(i2, i.cpu())


def test_knn_golden_cosine_and_errors(gold):
    from foundpose_b200.utils import knn_util

    repre, bank = _small_repre(gold)
    q = synthetic.make_query_features(150, 64, bank["feat_vectors"], seed=gold["knn/query_seed"])
    k3 = knn_util.KNN(k=3, metric="l2")
    k3.fit(bank["feat_cluster_centroids"].cuda())
    d, i = k3.search(q.cuda())
    assert torch.equal(i.cpu(), gold["knn/l2_k3_i"])
    assert torch.allclose(d.cpu(), gold["knn/l2_k3_d"], rtol=1e-3, atol=1e-4)
    kc = knn_util.KNN(k=2, metric="cosine")
    kc.fit(bank["feat_vectors"].cuda())
    dc, ic = kc.search(q.cuda())
    assert torch.equal(ic.cpu(), gold["knn/cos_k2_i"])
    assert torch.allclose(dc.cpu(), gold["knn/cos_k2_d"], rtol=0, atol=2e-3)   # fp16 unit vectors
    with pytest.raises(ValueError):
        knn_util.KNN(k=1, metric="l1").fit(q.cuda())


def test_tfidf_and_template_matching_vs_golden(gold):
    from foundpose_b200.utils import knn_util, template_util

    repre, bank = _small_repre(gold)
    q = synthetic.make_query_features(150, 64, bank["feat_vectors"], seed=gold["knn/query_seed"]).cuda()
    k3 = knn_util.KNN(k=3, metric="l2")
    k3.fit(bank["feat_cluster_centroids"].cuda())
    wid, wd = template_util.find_nearest_object_features(q, k3)
    assert torch.equal(wid.cpu(), gold["tfidf/word_ids"])
    idfs = gold["bank/idfs"].cuda()
    hard = template_util.calc_tfidf(wid, wd, idfs, soft_assignment=False)
    assert torch.allclose(hard.cpu(), gold["tfidf/hard"], rtol=1e-5, atol=1e-9)
    soft = template_util.calc_tfidf(wid, wd, idfs, soft_assignment=True, soft_sigma_squared=10.0)
    assert torch.allclose(soft.cpu(), gold["tfidf/soft"], rtol=2e-3, atol=1e-7)  # fp16 distance rounding
    ids, scores = template_util.template_matching(q, repre, 5, "tfidf", k3)
    assert torch.equal(ids.cpu(), gold["tfidf/top5_ids"])                    # bit-exact template ids
    assert torch.allclose(scores.cpu(), gold["tfidf/top5_scores"], rtol=0, atol=1e-5)
    with pytest.raises(ValueError):
        template_util.template_matching(q, repre, 5, "nope", k3)


def test_calc_tfidf_descriptors_vs_golden(gold):
    from foundpose_b200.utils import template_util

    repre, bank = _small_repre(gold)
    descs, idfs = template_util.calc_tfidf_descriptors(
        bank["feat_vectors"].cuda(), gold["bank/feat_to_word"].cuda(), bank["feat_to_template_ids"].cuda(),
        bank["feat_cluster_centroids"].cuda(), 24, 3, False, 10.0)
    assert torch.allclose(idfs.cpu(), gold["bank/idfs"], rtol=1e-6, atol=0)
    assert torch.allclose(descs.cpu(), gold["bank/template_descs"], rtol=1e-5, atol=1e-9)


def _check_corresp(ours, ref, k):
    assert len(ours) == len(ref)
    for a, b in zip(ours, ref):
        assert int(a["template_id"]) == int(b["template_id"])
        assert abs(float(a["template_score"]) - float(b["template_score"])) < 1e-5
        ad = a["nn_dists"].cpu() if "nn_dists" in a else None
        if ad is not None:
            assert torch.equal(ad, b["nn_dists"])
        assert torch.equal(a["coord_2d_ids"].cpu(), b["coord_2d_ids"])       # bit-exact 2D ids
        assert torch.equal(a["nn_vertex_ids"].cpu(), b["nn_vertex_ids"])     # bit-exact 3D ids
        assert torch.equal(a["coord_2d"].cpu(), b["coord_2d"])
        assert torch.equal(a["coord_3d"].cpu(), b["coord_3d"])
        assert torch.allclose(a["coord_conf"].cpu(), b["coord_conf"], equal_nan=True)


def test_establish_correspondences_vs_oracle_and_golden(gold):
    from foundpose_b200.utils import corresp_util, knn_util
    from oracle import corresp as ocorresp

    repre, bank = _small_repre(gold)
    q = synthetic.make_query_features(150, 64, bank["feat_vectors"], seed=gold["knn/query_seed"])
    grid = gold["corresp/grid"]
    k3 = knn_util.KNN(k=3, metric="l2")
    k3.fit(bank["feat_cluster_centroids"].cuda())
    ours = corresp_util.establish_correspondences(
        query_points=grid.cuda(), query_features=q.cuda(), object_repre=repre, template_matching_type="tfidf",
        feat_matching_type="cyclic_buddies", top_n_templates=5, top_k_buddies=40,
        visual_words_knn_index=k3, template_knn_indices=None, debug=True)
    bank_o = dict(bank)
    bank_o["template_descs"], bank_o["feat_cluster_idfs"] = gold["bank/template_descs"], gold["bank/idfs"]
    ref = ocorresp.establish_correspondences(grid, q, bank_o, 5, 40)
    _check_corresp(ours, ref, 40)
    # Against the reference's own output: same template ids / distance multiset (torch.topk leaves
    # the order of exact ties undefined there).
    for a, b in zip(ours, gold["corresp/list"]):
        assert int(a["template_id"]) == int(b["template_id"])
        assert torch.equal(a["nn_dists"].cpu(), b["nn_dists"])
    with pytest.raises(ValueError):
        corresp_util.establish_correspondences(grid.cuda(), q.cuda(), repre, "tfidf", "nope", 5, 40, k3)
    # cyclic_buddies_matching as a standalone call
    tpl0 = torch.nonzero(bank["feat_to_template_ids"] == 3).flatten()
    obj = bank["feat_vectors"][tpl0]
    qi = knn_util.KNN(1, "l2"); qi.fit(q.cuda())
    oi = knn_util.KNN(1, "l2"); oi.fit(obj.cuda())
    r = corresp_util.cyclic_buddies_matching(grid.cuda(), q.cuda(), qi, obj.cuda(), oi, 40, False)
    ro = ocorresp.cyclic_buddies_matching(grid, q, obj, 40)
    for x, y in zip(r, ro):
        assert torch.allclose(x.cpu().float(), y.float(), equal_nan=True)


def test_batched_pipeline_matches_per_crop_oracle():
    """B crops with ragged masks through RetrievalEngine == the oracle run crop by crop."""
    from foundpose_b200 import _native, pipeline
    from foundpose_b200.utils import feature_util, repre_util, template_util
    from oracle import corresp as ocorresp

    T, P, d, W, B = 40, 96, 128, 64, 6
    bank = synthetic.make_bank_tensors(T, P, d, num_words=W, seed=3, ragged=True)
    feat = bank["feat_vectors"]
    from oracle import knn as oknn, template as otemplate
    f2w = oknn.knn_l2(feat, bank["feat_cluster_centroids"], 1)[1].flatten()
    descs, idfs = otemplate.calc_tfidf_descriptors(feat, f2w, bank["feat_to_template_ids"],
                                                   bank["feat_cluster_centroids"], T, 3, False, 10.0)
    repre = repre_util.FeatureBasedObjectRepre(
        vertices=bank["vertices"], feat_vectors=feat, feat_to_template_ids=bank["feat_to_template_ids"],
        feat_cluster_centroids=bank["feat_cluster_centroids"], feat_cluster_idfs=idfs, template_descs=descs,
        template_desc_opts=repre_util.TemplateDescOpts())
    index = pipeline.ObjectIndex(repre, torch.device("cuda"))
    grid = feature_util.generate_grid_points((140, 140), 14.0)     # 100 points per crop
    stride = grid.shape[0]
    masks = synthetic.make_masks(B, (140, 140), seed=9)
    masks[2] = False                                                # an empty crop
    pts = torch.zeros(B, stride, 2, device="cuda")
    ids = torch.zeros(B, stride, dtype=torch.int32, device="cuda")
    cnt = torch.zeros(B, dtype=torch.int32, device="cuda")
    _native.filter_points_by_mask(grid.cuda().contiguous(), masks.to(torch.uint8).cuda().contiguous(), pts, ids, cnt)
    counts = cnt.cpu().tolist()
    feats = torch.zeros(B * stride, d)
    per_crop = []
    for b in range(B):
        qb = synthetic.make_query_features(counts[b], d, feat, seed=100 + b)
        feats[b * stride: b * stride + counts[b]] = qb
        per_crop.append(qb)
    engine = pipeline.RetrievalEngine(index, B, stride, 5, 30)
    out = engine.match(feats.half().cuda().contiguous(), pts, cnt)
    torch.cuda.synchronize()
    bank_o = dict(bank); bank_o["template_descs"], bank_o["feat_cluster_idfs"] = descs, idfs
    for b in range(B):
        ours = pipeline.outputs_to_corresp_list(out, b, debug=True)
        if counts[b] == 0:
            assert all(len(c["coord_2d"]) == 0 for c in ours)
            continue
        ref = ocorresp.establish_correspondences(pts[b, :counts[b]].cpu(), per_crop[b], bank_o, 5, 30)
        _check_corresp(ours, ref, 30)


@pytest.mark.parametrize("nq,top_k", [(6000, 300), (40000, 300), (9000, 2048)])
def test_cyclic_buddies_large_query_sets(nq, top_k):
    """More than 4096 query points per crop (the reference's default grid_cell_size = 1 gives 176 400):
    chunk-wise pre-selection + final sort must equal the oracle's canonical top-k."""
    from foundpose_b200.utils import corresp_util, knn_util
    from oracle import corresp as ocorresp

    g = torch.Generator().manual_seed(nq)
    obj = synthetic.fp16_representable(torch.randn(700, 64, generator=g))
    q = synthetic.make_query_features(nq, 64, obj, seed=3, noise=0.5)
    # points on a coarse lattice -> many exact ties in the cyclic distance
    pts = torch.stack([torch.randint(0, 60, (nq,), generator=g), torch.randint(0, 60, (nq,), generator=g)], 1).float() * 7.0
    qi = knn_util.KNN(1, "l2"); qi.fit(q.cuda())
    oi = knn_util.KNN(1, "l2"); oi.fit(obj.cuda())
    r = corresp_util.cyclic_buddies_matching(pts.cuda(), q.cuda(), qi, obj.cuda(), oi, top_k, False)
    ro = ocorresp.cyclic_buddies_matching(pts, q, obj, top_k)
    assert torch.equal(r[0].cpu(), ro[0])            # query ids, canonical tie order
    assert torch.equal(r[1].cpu(), ro[1])            # object ids
    assert torch.equal(r[2].cpu(), ro[2])            # distances
    assert torch.allclose(r[3].cpu(), ro[3], equal_nan=True)


@pytest.mark.parametrize("soft", [False, True])
def test_engine_cosine_visual_word_metric(soft):
    """tfidf_knn_metric="cosine" through the batched engine == oracle tfidf_matching(knn_metric="cosine")."""
    from foundpose_b200 import pipeline
    from foundpose_b200.utils import repre_util
    from oracle import knn as oknn, template as otemplate

    T, P, d, W, B, nq = 30, 64, 128, 48, 3, 80
    bank = synthetic.make_bank_tensors(T, P, d, num_words=W, seed=5)
    feat = bank["feat_vectors"]
    f2w = oknn.knn_l2(feat, bank["feat_cluster_centroids"], 1)[1].flatten()
    descs, idfs = otemplate.calc_tfidf_descriptors(feat, f2w, bank["feat_to_template_ids"],
                                                   bank["feat_cluster_centroids"], T, 3, False, 10.0)
    opts = repre_util.TemplateDescOpts(tfidf_knn_metric="cosine", tfidf_soft_assign=soft, tfidf_soft_sigma_squared=0.05)
    repre = repre_util.FeatureBasedObjectRepre(
        vertices=bank["vertices"], feat_vectors=feat, feat_to_template_ids=bank["feat_to_template_ids"],
        feat_cluster_centroids=bank["feat_cluster_centroids"], feat_cluster_idfs=idfs, template_descs=descs,
        template_desc_opts=opts)
    index = pipeline.ObjectIndex(repre, torch.device("cuda"))
    engine = pipeline.RetrievalEngine(index, B, nq, 5, 20)
    engine.full_scores = True          # also produce the [B, T] score matrix (fp32 path) for the comparison below
    qs = [synthetic.make_query_features(nq, d, feat, seed=40 + b) for b in range(B)]
    pts = torch.rand(B, nq, 2, device="cuda") * 100
    cnt = torch.full((B,), nq, dtype=torch.int32, device="cuda")
    out = engine.match(torch.cat(qs).half().cuda().contiguous(), pts, cnt)
    torch.cuda.synchronize()
    for b in range(B):
        ids, scores, tfidf, cos = otemplate.tfidf_matching(qs[b].half().float(), bank["feat_cluster_centroids"], idfs,
                                                           descs, 5, 3, "cosine", soft, 0.05)
        assert torch.allclose(out.query_tfidf[b].cpu(), tfidf, rtol=5e-3, atol=1e-6)
        assert torch.allclose(out.cos_sims[b].cpu(), cos, rtol=0, atol=2e-4)
        gap = (cos[ids][:-1] - cos[ids][1:]).min()
        if gap > 1e-3:
            assert torch.equal(out.template_ids[b].cpu(), ids)


def test_pair_kernel_sweep_barrier_gives_up_instead_of_hanging():
    """The sweep barrier of fp_knn_search_pair_items is an optimisation only: items that announce more participants
    than there are clusters (a barrier that can never be satisfied - what a grid that is not co-resident looks like)
    must switch it off after a bounded wait and still return the exact result."""
    from foundpose_b200 import _native
    from foundpose_b200.utils import knn_util
    from oracle import knn as oknn

    g = torch.Generator().manual_seed(12)
    nq, nb, dim, k = 74 * 256, 9000, 64, 5
    bank = synthetic.fp16_representable(torch.randn(nb, dim, generator=g))
    q = synthetic.make_query_features(nq, dim, bank, seed=6)
    bank16, q16 = bank.half().cuda(), q.half().cuda()
    bn, qn = _native.row_sqnorm_f16(bank16), _native.row_sqnorm_f16(q16)
    direct, split, chunks, _ = knn_util.plan_pair_items(nq, nb, _native.num_sms() // 2)
    assert not split
    items = _native.knn_items_from_host([it + (len(direct) + 1,) for it in direct], "cuda")   # one participant too many
    d = torch.empty((nq, k), device="cuda")
    i = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    sync = torch.zeros(2, dtype=torch.int64, device="cuda")
    _native.knn_search_pair_items(q16, qn, bank16, bn, items, len(direct), 0, k, d, i, sync, 4)
    torch.cuda.synchronize()
    assert int(sync[1].item()) == 1                               # the give-up flag was raised
    rd, ri = oknn.knn_l2(q[:2048], bank, k)
    sure = oknn.topk_margin(q[:2048], bank, k) > 1e-4
    assert torch.equal(i[:2048].cpu()[sure], ri[sure])
    # and with the right participant count the flag stays down
    items = _native.knn_items_from_host([it + (len(direct),) for it in direct], "cuda")
    _native.knn_search_pair_items(q16, qn, bank16, bn, items, len(direct), 0, k, d, i, sync, 4)
    torch.cuda.synchronize()
    assert int(sync[1].item()) == 0 and int(sync[0].item()) > 0
    assert torch.equal(i[:2048].cpu()[sure], ri[sure])
