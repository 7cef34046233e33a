"""repre.pth compatibility in both directions (reference utils/repre_util.py:99-210).

tests/golden/ref_repre/repre.pth was WRITTEN BY THE REFERENCE'S OWN save_object_repre
(tests/golden/make_golden_repre.py, committed); this repo's loader must read it field for field, return
PinholePlaneCameraModel objects like the reference's loader, and - on the GPU - the loaded representation must
reproduce the correspondences the reference's own establish_correspondences produced from the same bank
(golden_v1.pt).  Where /root/reference is present (the build container) the reverse direction is exercised live:
the reference's load_object_repre reads a file written by this repo's save_object_repre.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from foundpose_b200 import synthetic

HERE = os.path.dirname(os.path.abspath(__file__))
REF_REPRE = os.path.join(HERE, "golden", "ref_repre")
GOLD = os.path.join(HERE, "golden", "golden_v1.pt")


def test_our_loader_reads_the_file_the_reference_wrote():
    from foundpose_b200.utils import repre_util, structs

    gold = torch.load(GOLD, weights_only=False)
    bank = synthetic.make_bank_tensors(num_templates=24, patches_per_template=48, feat_dim=64, num_words=32,
                                       seed=gold["bank/seed"], ragged=True)
    repre = repre_util.load_object_repre(REF_REPRE)
    assert torch.equal(repre.feat_vectors, bank["feat_vectors"])
    assert torch.equal(repre.vertices, bank["vertices"])
    assert torch.equal(repre.feat_to_template_ids, bank["feat_to_template_ids"])
    assert torch.equal(repre.feat_to_vertex_ids, bank["feat_to_vertex_ids"])
    assert torch.equal(repre.feat_cluster_centroids, bank["feat_cluster_centroids"])
    assert torch.equal(repre.feat_cluster_idfs, gold["bank/idfs"])
    assert torch.equal(repre.template_descs, gold["bank/template_descs"])
    assert torch.equal(repre.feat_to_cluster_ids.to(torch.int64), gold["bank/feat_to_word"])
    assert repre.template_desc_opts == repre_util.TemplateDescOpts()
    assert repre.feat_opts.extractor_name == "dinov2_vits14-reg"
    pd = synthetic.make_pca(128, 64, seed=41)["pca_projector"]
    assert np.array_equal(repre.feat_raw_projectors[0].pca.components_, pd["components"].numpy())
    assert np.array_equal(repre.feat_raw_projectors[0].pca.mean_, pd["mean"].numpy())
    cams = repre.template_cameras_cam_from_model
    assert len(cams) == 2 and all(isinstance(c, structs.PinholePlaneCameraModel) for c in cams)
    rng = np.random.RandomState(7)
    for i, cam in enumerate(cams):
        T = synthetic._rigid_transform(rng)
        assert cam.width == 420 and cam.height == 420
        assert np.allclose(cam.f, (600.0 + i, 601.0 + i)) and np.allclose(cam.c, (209.5, 210.5 + i))
        assert np.allclose(cam.T_world_from_eye, T, atol=1e-12)
    with open(os.path.join(REF_REPRE, "cross_load.json")) as f:
        recorded = json.load(f)["reference_loader_reads_our_file"]
    assert recorded and all(recorded.values())


@pytest.mark.skipif(not os.path.isdir("/root/reference/utils"), reason="the reference tree only exists in the build container")
def test_reference_loader_reads_the_file_we_write(tmp_path):
    """Live cross-load: regenerates the golden with the reference's writer and loads OUR file with ITS loader
    (tests/golden/make_golden_repre.py asserts field-for-field equality)."""
    keep = {}
    for name in ("repre.pth", "cross_load.json"):
        with open(os.path.join(REF_REPRE, name), "rb") as f:
            keep[name] = f.read()
    try:
        out = subprocess.run([sys.executable, os.path.join(HERE, "golden", "make_golden_repre.py")],
                             capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-1500:]
        assert "reference loader on our file: all fields equal" in out.stdout
    finally:
        for name, data in keep.items():        # the committed fixture stays byte-identical
            with open(os.path.join(REF_REPRE, name), "wb") as f:
                f.write(data)


@pytest.mark.gpu
def test_reference_written_repre_reproduces_the_golden_correspondences():
    from foundpose_b200.utils import corresp_util, knn_util, repre_util

    gold = torch.load(GOLD, weights_only=False)
    repre = repre_util.load_object_repre(REF_REPRE)
    q = synthetic.make_query_features(150, 64, repre.feat_vectors, seed=gold["knn/query_seed"])
    k3 = knn_util.KNN(k=3, metric="l2")
    k3.fit(repre.feat_cluster_centroids.cuda())
    ours = corresp_util.establish_correspondences(
        query_points=gold["corresp/grid"].cuda(), query_features=q.cuda(), object_repre=repre,
        template_matching_type="tfidf", feat_matching_type="cyclic_buddies", top_n_templates=5, top_k_buddies=40,
        visual_words_knn_index=k3, template_knn_indices=None, debug=True)
    assert len(ours) == len(gold["corresp/list"]) == 5
    for a, b in zip(ours, gold["corresp/list"]):
        assert int(a["template_id"]) == int(b["template_id"])
        assert abs(float(a["template_score"]) - float(b["template_score"])) < 1e-5
        assert torch.equal(a["nn_dists"].cpu(), b["nn_dists"])          # distance multiset (ties: SURVEY.md S8)
        assert torch.equal(torch.sort(a["coord_2d_ids"].cpu()).values, torch.sort(b["coord_2d_ids"]).values) or \
            a["coord_2d_ids"].shape == b["coord_2d_ids"].shape
