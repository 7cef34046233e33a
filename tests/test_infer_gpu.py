"""The infer.py entry point on the GPU: the reference-style per-crop loop and the batched pipeline
must produce the same correspondences; unsorted template ids go through the feature permutation."""
import pytest
import torch

from foundpose_b200 import synthetic

pytestmark = pytest.mark.gpu


def test_per_crop_and_batched_modes_agree(monkeypatch):
    from foundpose_b200.scripts import infer

    monkeypatch.setenv("FOUNDPOSE_SYNTHETIC_WEIGHTS", "5")
    opts = infer.InferOpts(extractor_name="dinov2_version=tiny-test-reg_stride=14_facet=token_layer=2_norm=1",
                           grid_cell_size=14.0, crop_size=(112, 112), match_top_n_templates=3,
                           match_top_k_buddies=25, debug=False)
    a = infer.infer(opts, num_synthetic_crops=5, batch=0, synthetic_bank=(12, 40, 128, 32))
    b = infer.infer(opts, num_synthetic_crops=5, batch=4, synthetic_bank=(12, 40, 128, 32))
    assert len(a) == len(b) == 5
    for ra, rb in zip(a, b):
        assert ra["crop_id"] == rb["crop_id"]
        # coarse-pose block runs in both modes (synthetic correspondences carry no geometry: a pose may be None)
        assert "best_coarse_pose" in ra and "best_coarse_pose" in rb
        for r in (ra, rb):
            if r["best_coarse_pose"] is not None:
                assert r["best_coarse_pose"]["R_m2c"].shape == (3, 3) and r["best_coarse_pose"]["quality"] >= 6
        if len(ra["corresp"]) == 0:          # empty mask -> the per-crop path returns no correspondences
            assert all(len(c["coord_2d"]) == 0 for c in rb["corresp"])
            continue
        for ca, cb in zip(ra["corresp"], rb["corresp"]):
            assert int(ca["template_id"]) == int(cb["template_id"])
            assert torch.equal(ca["coord_2d_ids"], cb["coord_2d_ids"])
            assert torch.equal(ca["nn_vertex_ids"], cb["nn_vertex_ids"])
            assert torch.allclose(ca["coord_3d"], cb["coord_3d"])


def test_unsorted_template_ids_use_the_permutation():
    from foundpose_b200 import pipeline
    from foundpose_b200.utils import corresp_util, knn_util, repre_util, template_util
    from oracle import corresp as ocorresp

    bank = synthetic.make_bank_tensors(10, 30, 64, num_words=64, seed=4, ragged=True)
    g = torch.Generator().manual_seed(0)
    perm = torch.randperm(bank["feat_vectors"].shape[0], generator=g)
    feat, tpl, verts = bank["feat_vectors"][perm], bank["feat_to_template_ids"][perm], bank["vertices"][perm]
    wk = knn_util.KNN(1, "l2"); wk.fit(bank["feat_cluster_centroids"].cuda())
    f2w = wk.search(feat.cuda())[1].flatten()
    descs, idfs = template_util.calc_tfidf_descriptors(feat.cuda(), f2w, tpl.cuda(), bank["feat_cluster_centroids"].cuda(),
                                                       10, 3, False, 10.0)
    repre = repre_util.FeatureBasedObjectRepre(
        vertices=verts, feat_vectors=feat, feat_to_template_ids=tpl, feat_cluster_centroids=bank["feat_cluster_centroids"],
        feat_cluster_idfs=idfs.cpu(), template_descs=descs.cpu(), template_desc_opts=repre_util.TemplateDescOpts())
    index = pipeline.get_object_index(repre, torch.device("cuda"))
    assert index.feat_perm is not None
    q = synthetic.make_query_features(60, 64, feat, seed=9)
    pts = torch.rand(60, 2, generator=g) * 100
    k3 = knn_util.KNN(3, "l2"); k3.fit(bank["feat_cluster_centroids"].cuda())   # k comes from the index, as in the reference
    ours = corresp_util.establish_correspondences(pts.cuda(), q.cuda(), repre, "tfidf", "cyclic_buddies", 3, 20, k3, None, True)
    ref = ocorresp.establish_correspondences(pts, q, {
        "feat_vectors": feat, "feat_to_template_ids": tpl, "vertices": verts,
        "feat_cluster_centroids": bank["feat_cluster_centroids"], "feat_cluster_idfs": idfs.cpu(),
        "template_descs": descs.cpu()}, 3, 20)
    from oracle import template as otemplate

    _, _, _, cos = otemplate.tfidf_matching(q, bank["feat_cluster_centroids"], idfs.cpu(), descs.cpu(), 3)
    top = torch.sort(cos, descending=True).values
    assert float((top[:3] - top[1:4]).min()) > 1e-5, "test data has a near-tie in the retrieval scores"
    for a, b in zip(ours, ref):
        assert int(a["template_id"]) == int(b["template_id"])
        assert torch.equal(a["nn_vertex_ids"].cpu(), b["nn_vertex_ids"])      # original feature ids
        assert torch.equal(a["coord_3d"].cpu(), b["coord_3d"])


def test_real_crops_use_their_own_cameras_and_need_them_for_poses(tmp_path, monkeypatch):
    """--crops blobs: coarse poses are computed in the per-crop camera the blob brings (`intrinsics` [P,4]); a blob
    without cameras yields correspondences but no poses (no made-up camera for real data)."""
    from foundpose_b200.scripts import infer

    monkeypatch.setenv("FOUNDPOSE_SYNTHETIC_WEIGHTS", "5")
    opts = infer.InferOpts(extractor_name="dinov2_version=tiny-test-reg_stride=14_facet=token_layer=2_norm=1",
                           grid_cell_size=14.0, crop_size=(112, 112), match_top_n_templates=3,
                           match_top_k_buddies=25, debug=False)
    images = synthetic.make_crops(4, (112, 112), seed=1)
    masks = synthetic.make_masks(4, (112, 112), seed=2)
    with_cam, without_cam = str(tmp_path / "a.pt"), str(tmp_path / "b.pt")
    intr = torch.tensor([[500.0 + i, 510.0 + i, 56.0, 57.0] for i in range(4)], dtype=torch.float64)
    torch.save({"images": images, "masks": masks, "intrinsics": intr}, with_cam)
    torch.save({"images": images, "masks": masks}, without_cam)
    for batch in (0, 4):
        res = infer.infer(opts, crops_path=with_cam, batch=batch, synthetic_bank=(12, 40, 128, 32))
        assert len(res) == 4 and all("best_coarse_pose" in r for r in res)
        res = infer.infer(opts, crops_path=without_cam, batch=batch, synthetic_bank=(12, 40, 128, 32))
        assert len(res) == 4 and all(r["best_coarse_pose"] is None for r in res)
        assert any(len(r["corresp"]) > 0 for r in res)
    with pytest.raises(ValueError):
        infer.infer(infer.InferOpts(**{**opts._asdict(), "max_num_queries": 10}), crops_path=with_cam, batch=4,
                    synthetic_bank=(12, 40, 128, 32))
