"""CPU-side checks: the C ABI library loads and exports every declared symbol (no compute calls),
and the host-side logic of the reference-API mirror (name grammar, options, repre.pth round trip,
sharding) behaves like the reference."""
import ctypes
import json
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from foundpose_b200 import _native

    declared = _native.declared_symbols()
    assert len(declared) >= 28 and "fp_vit_forward" in declared and "fp_knn_search_items" in declared
    lib = _native.load()          # raises ImportError when a declared symbol is missing
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.fp_version() >= 100
    assert lib.fp_last_error() is not None
    # every exported fp_* symbol is declared in the header (no undocumented entry points)
    out = subprocess.run(["nm", "-D", "--defined-only", _native.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (fp_[a-z0-9_]+)", out)))
    assert exported == declared


def test_knn_item_struct_layout_matches_header():
    from foundpose_b200 import _native

    class Item(ctypes.Structure):
        _fields_ = [("q_row0", ctypes.c_int32), ("q_rows", ctypes.c_int32), ("b_row0", ctypes.c_int32),
                    ("b_rows", ctypes.c_int32), ("out_row0", ctypes.c_int64), ("reserved", ctypes.c_int64)]

    assert ctypes.sizeof(Item) == _native.KNN_ITEM_BYTES == 32


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "foundpose_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), os.path.join(dirpath, f)


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from foundpose_b200 import _native

    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(ImportError):
        _native.load()


def test_extractor_name_grammar_matches_reference():
    from foundpose_b200.utils import dinov2_utils

    o = dinov2_utils.parse_model_name("dinov2_vitl14")
    assert (o["version"], o["stride"], o["facet"], o["layer"], o["norm"]) == ("vitl14", 14, "token", 9, True)
    o = dinov2_utils.parse_model_name("dinov2_version=vits14-reg_stride=14_facet=token_layer=9_logbin=0_norm=1")
    assert (o["version"], o["layer"], o["norm"]) == ("vits14-reg", 9, True)   # unknown keys are ignored
    o = dinov2_utils.parse_model_name("dinov2_version=vitl14_stride=14_facet=key_layer=18_norm=0")
    assert (o["facet"], o["layer"], o["norm"]) == ("key", 18, False)
    with pytest.raises(AssertionError):
        dinov2_utils.parse_model_name("dino_vitl14")
    with pytest.raises(ValueError):
        dinov2_utils.parse_model_name("dinov2_vits14_reg")
    from foundpose_b200.utils import feature_util

    with pytest.raises(NotImplementedError):
        feature_util.make_feature_extractor("resnet50")


def test_weights_container_uses_dinov2_state_dict_layout():
    from foundpose_b200 import synthetic
    from foundpose_b200.utils import dinov2_utils

    arch = synthetic.VIT_ARCHS["tiny-test-reg"]
    sd = synthetic.make_vit_state_dict(arch, seed=3)
    ext = dinov2_utils.DinoFeatureExtractor("dinov2_version=tiny-test-reg_layer=1", state_dict=sd)
    got = ext.model.state_dict()
    assert set(got) == set(sd)
    for k in sd:
        assert torch.equal(got[k], sd[k])
    # an official-style checkpoint loads with strict=True
    ext.model.load_state_dict(synthetic.make_vit_state_dict(arch, seed=4), strict=True)
    with pytest.raises(ValueError):
        ext(torch.rand(1, 3, 56, 56))          # CPU tensors are rejected: no CPU fallback
    with pytest.raises(NotImplementedError):
        dinov2_utils.DinoFeatureExtractor("dinov2_version=tiny-test_stride=7", state_dict=sd)


def test_infer_opts_read_the_reference_config_format(tmp_path):
    from foundpose_b200.scripts import infer

    cfg = {"infer_opts": {"version": "v1", "object_dataset": "lmo", "repre_version": "v1", "crop_rel_pad": 0.2,
                          "crop_size": [420, 420], "use_detections": True,
                          "extractor_name": "dinov2_version=vits14-reg_stride=14_facet=token_layer=9_logbin=0_norm=1",
                          "grid_cell_size": 14.0, "match_template_type": "tfidf", "match_top_n_templates": 5,
                          "match_feat_matching_type": "cyclic_buddies", "match_top_k_buddies": 300,
                          "pnp_type": "opencv", "pnp_ransac_iter": 400, "pnp_inlier_thresh": 10.0,
                          "final_pose_type": "best_coarse", "num_preds_factor": 1, "vis_results": True}}
    p = tmp_path / "lmo.json"
    p.write_text(json.dumps(cfg))
    opts = infer.load_opts(str(p), {})
    assert opts.grid_cell_size == 14.0 and opts.crop_size == (420, 420) and opts.match_top_k_buddies == 300
    assert infer.InferOpts().extractor_name == "dinov2_vitl14" and infer.InferOpts().grid_cell_size == 1.0
    with pytest.raises(ValueError):
        infer.load_opts(None, {"not_a_field": 1})


def test_repre_pth_round_trip(tmp_path):
    from foundpose_b200 import synthetic
    from foundpose_b200.utils import projector_util, repre_util

    bank = synthetic.make_bank_tensors(6, 10, 64, num_words=8, seed=1, ragged=True)
    repre = repre_util.FeatureBasedObjectRepre(
        vertices=bank["vertices"], feat_vectors=bank["feat_vectors"],
        feat_to_template_ids=bank["feat_to_template_ids"], feat_cluster_centroids=bank["feat_cluster_centroids"],
        feat_cluster_idfs=torch.rand(8), template_descs=torch.rand(6, 8),
        template_desc_opts=repre_util.TemplateDescOpts(), feat_opts=repre_util.FeatureOpts("dinov2_vits14-reg"),
        feat_raw_projectors=[projector_util.projector_from_tensordict(synthetic.make_pca(128, 64, 2))],
        template_cameras_cam_from_model=[{"f": torch.ones(2), "c": torch.zeros(2), "width": 420, "height": 420,
                                          "T_world_from_eye": torch.eye(4)}])
    repre_util.save_object_repre(repre, str(tmp_path))
    raw = torch.load(str(tmp_path / "repre.pth"), weights_only=False)
    # same dictionary layout as the reference writes (utils/repre_util.py:99-141)
    assert {"feat_vectors", "feat_opts", "template_desc_opts", "feat_raw_projectors", "feat_vis_projectors",
            "template_cameras_cam_from_model"} <= set(raw)
    assert set(raw["feat_raw_projectors"][0]["pca_projector"]) == {
        "components", "explained_variance", "explained_variance_ratio", "singular_values", "mean",
        "noise_variance", "whiten"}
    back = repre_util.load_object_repre(str(tmp_path))
    assert torch.equal(back.feat_vectors, repre.feat_vectors)
    assert torch.equal(back.feat_to_template_ids, repre.feat_to_template_ids)
    assert back.template_desc_opts == repre_util.TemplateDescOpts() and back.feat_opts.extractor_name == "dinov2_vits14-reg"
    assert back.feat_raw_projectors[0].pca.components_.shape == (64, 128)
    assert repre_util.get_object_repre_dir_path("a", "v1", "lmo", 5) == os.path.join("a", "lmo", "v1", "5")
    as_np = repre_util.convert_object_repre_to_numpy(back)
    assert as_np.feat_vectors.shape == (bank["feat_vectors"].shape[0], 64)


def test_shard_helpers_cover_all_items_exactly_once():
    from foundpose_b200 import distributed

    for n in (0, 1, 7, 64, 4096):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                s, e = distributed.shard_range(n, r, world)
                assert 0 <= s <= e <= n
                seen += list(range(s, e))
            assert seen == list(range(n))
            rr = sorted(i for r in range(world) for i in distributed.shard_round_robin(n, r, world))
            assert rr == list(range(n))


def test_cpu_inputs_are_rejected_by_every_entry_point():
    from foundpose_b200.utils import corresp_util, feature_util, projector_util, template_util
    from foundpose_b200 import synthetic

    with pytest.raises(ValueError):
        feature_util.filter_points_by_mask(torch.zeros(4, 2), torch.ones(8, 8))
    with pytest.raises(ValueError):
        feature_util.sample_feature_map_at_points(torch.zeros(8, 4, 4), torch.zeros(2, 2), (56, 56))
    proj = projector_util.projector_from_tensordict(synthetic.make_pca(128, 64, 0))
    with pytest.raises(ValueError):
        proj.transform(torch.zeros(3, 128))
    with pytest.raises(ValueError):
        template_util.calc_tfidf(torch.zeros(3, 3, dtype=torch.int64), torch.zeros(3, 3), torch.ones(8), False)


def test_knn_split_chunk_count_fills_whole_waves():
    """Host logic of the split k-NN path: the fewest bank chunks whose items fill whole waves of the 148 CTAs."""
    import math

    from foundpose_b200.utils import knn_util

    assert knn_util._choose_num_chunks(1) in range(141, 149)          # one chunk per SM for a single query block
    for n_q in (1, 2, 3, 8, 37, 57, 73):
        c = knn_util._choose_num_chunks(n_q)
        waves = n_q * c / 148
        assert 1 <= c <= max(1, 592 // n_q)
        assert waves / math.ceil(waves) >= 0.93, (n_q, c)


def test_pair_item_plan_covers_every_query_once_and_fills_the_tail_wave():
    """Host logic of the tensor-bound k-NN path (fp_knn_search_pair_items): whole waves of 74 CTA pairs sweep the
    whole bank, the items of the last partial wave are split into bank slices that together cover it exactly."""
    from foundpose_b200.utils import knn_util

    for nq, nb in [(460800, 10240000), (57600, 10240000), (21541, 9000), (9473, 8193), (256 * 74, 50000), (300, 9000)]:
        direct, split, chunks, chunk_rows = knn_util.plan_pair_items(nq, nb, 74)
        assert len(direct) % 74 == 0 or not split     # a tail of more than 37 items cannot be split: left whole
        covered = sorted((q0, q0 + qr) for q0, qr, b0, br, _ in direct)
        tail = {}
        for q0, qr, b0, br, o0 in split:
            assert 0 < br <= chunk_rows and b0 % 256 == 0 and chunk_rows % 256 == 0
            tail.setdefault((q0, qr), []).append((b0, br))
        for (q0, qr), slices in tail.items():
            slices.sort()
            assert slices[0][0] == 0 and sum(br for _, br in slices) == nb            # slices tile the bank
            assert all(a[0] + a[1] == b[0] for a, b in zip(slices, slices[1:]))
            assert len(slices) == chunks
            covered.append((q0, q0 + qr))
        covered.sort()
        assert covered[0][0] == 0 and covered[-1][1] == nq
        assert all(a[1] == b[0] for a, b in zip(covered, covered[1:]))               # every query exactly once
        if split:
            assert len(split) <= 74 and len(split) > 74 // 2                          # the tail wave is (nearly) full
        # partial results of slice c, tail item j live at rows c * q_pad + j * 256 of the partial buffer
        q_pad = (len(split) // chunks) * 256 if split else 0
        assert sorted(o0 for *_, o0 in split) == sorted(c * q_pad + j * 256 for c in range(chunks if split else 0)
                                                        for j in range(len(split) // chunks))


def test_projector_chain_composes_to_one_affine_map():
    """project_features applies the projectors in turn (reference utils/projector_util.py:71-88); the batched pipeline
    folds the chain into ONE PCA-shaped projector: same result to fp32 rounding."""
    import numpy as np

    from foundpose_b200 import synthetic
    from foundpose_b200.utils import projector_util as pu

    chain = [pu.projector_from_tensordict(synthetic.make_pca(128, 64, 1)),
             pu.projector_from_tensordict(synthetic.make_pca(64, 32, 2)),
             pu.projector_from_tensordict(synthetic.make_pca(32, 16, 3))]
    one = pu.compose_projectors(chain)
    assert one.pca.components_.shape == (16, 128)
    x = np.random.RandomState(0).randn(40, 128)

    def transform(p, v):
        return (v - np.asarray(p.pca.mean_, dtype=np.float64)) @ np.asarray(p.pca.components_, dtype=np.float64).T

    ref = x
    for p in chain:
        ref = transform(p, ref)
    assert np.abs(transform(one, x) - ref).max() < 1e-6
    assert pu.compose_projectors(chain[:1]) is chain[0]
