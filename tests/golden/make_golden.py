"""Generates tests/golden/golden_v1.pt by running the REFERENCE'S OWN modules on seeded inputs.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

What runs unmodified from /root/reference: utils/dinov2_utils.py (DinoFeatureExtractor),
external/dinov2 (hub backbones, DinoVisionTransformer), utils/feature_util.py,
utils/projector_util.py (+ scikit-learn PCA), utils/knn_util.py, utils/template_util.py,
utils/corresp_util.py, utils/repre_util.py.

Stubs injected for modules that are not installed in this image:
  * faiss / faiss.contrib.torch_utils - IndexFlatL2 / IndexFlatIP restating faiss 1.8.0's
    exhaustive fp32 search (the arithmetic is shared with oracle/knn.py, so the k-NN leg is pinned
    to that published algorithm, not to faiss binaries)
  * kornia, torchinfo - empty modules (imported by the reference, unused on this path)
The hub factories are called with pretrained=False (no network) and receive the seeded weights of
foundpose_b200.synthetic.make_vit_state_dict via load_state_dict(strict=True).
"""

import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "external", "dinov2"))

from oracle import knn as oknn  # noqa: E402
from foundpose_b200 import synthetic  # noqa: E402


def install_stubs() -> None:
    faiss = types.ModuleType("faiss")

    class _Flat:
        def __init__(self, d):
            self.d = d
            self.rows = torch.empty(0, d)

        def train(self, x):
            pass

        def add(self, x):
            assert x.dtype == torch.float32 and x.is_contiguous()
            self.rows = torch.cat([self.rows, x.detach().clone()])

    class IndexFlatL2(_Flat):
        def search(self, x, k):
            assert x.dtype == torch.float32 and x.is_contiguous()
            return oknn.knn_l2(x, self.rows, k)

    class IndexFlatIP(_Flat):
        def search(self, x, k):
            sim = x.to(torch.float32) @ self.rows.t()
            order = torch.sort(-sim, dim=1, stable=True)
            return -order.values[:, :k].contiguous(), order.indices[:, :k].contiguous()

    faiss.IndexFlatL2 = IndexFlatL2
    faiss.IndexFlatIP = IndexFlatIP
    contrib = types.ModuleType("faiss.contrib")
    torch_utils = types.ModuleType("faiss.contrib.torch_utils")
    faiss.contrib = contrib
    contrib.torch_utils = torch_utils
    sys.modules["faiss"] = faiss
    sys.modules["faiss.contrib"] = contrib
    sys.modules["faiss.contrib.torch_utils"] = torch_utils
    for name in ("kornia", "torchinfo"):
        m = types.ModuleType(name)
        if name == "torchinfo":
            m.summary = lambda *a, **k: None
        sys.modules[name] = m


def build_reference_extractor(model_name: str, sd):
    """DinoFeatureExtractor with pretrained=False and the seeded synthetic weights."""
    import dinov2.hub.backbones as backbones
    from utils import dinov2_utils

    originals = {}
    for fn_name in list(backbones.__dict__):
        if fn_name.startswith("dinov2_vit"):
            orig = backbones.__dict__[fn_name]
            originals[fn_name] = orig

            def wrapped(*, pretrained=True, _orig=orig, **kw):
                return _orig(pretrained=False, **kw)

            backbones.__dict__[fn_name] = wrapped
    try:
        ext = dinov2_utils.DinoFeatureExtractor(model_name=model_name)
    finally:
        for k, v in originals.items():
            backbones.__dict__[k] = v
    ext.model.load_state_dict(sd, strict=True)
    return ext


def build_tiny_reference(arch, sd, layer: int):
    """A DinoFeatureExtractor whose hub model is replaced by a tiny DinoVisionTransformer."""
    from functools import partial

    from dinov2.layers import MemEffAttention
    from dinov2.layers import NestedTensorBlock as Block
    from dinov2.models.vision_transformer import DinoVisionTransformer
    from utils import dinov2_utils

    model = DinoVisionTransformer(
        img_size=arch.img_size, patch_size=arch.patch_size, embed_dim=arch.embed_dim, depth=arch.depth,
        num_heads=arch.num_heads, mlp_ratio=arch.mlp_ratio, init_values=1.0, block_chunks=0,
        block_fn=partial(Block, attn_class=MemEffAttention),
        num_register_tokens=arch.num_register_tokens,
        interpolate_antialias=arch.interpolate_antialias, interpolate_offset=arch.interpolate_offset)
    model.load_state_dict(sd, strict=True)
    model.eval()
    ext = dinov2_utils.DinoFeatureExtractor.__new__(dinov2_utils.DinoFeatureExtractor)
    torch.nn.Module.__init__(ext)
    import torchvision.transforms as T

    ext.version, ext.stride, ext.facet, ext.layer, ext.apply_norm = arch.name, 14, "token", layer, True
    ext.model = model
    ext.patch_size = 14
    ext._feats, ext.hook_handlers, ext.num_patches = [], [], None
    ext.normalize = T.Normalize(mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225))
    return ext


def main() -> None:
    install_stubs()
    from utils import corresp_util, feature_util, knn_util, projector_util, repre_util, template_util

    gold = {}

    # ---- A2-A4: ViT feature extraction --------------------------------------------------------
    for arch_name, layer, size in [("tiny-test", 1, 56), ("tiny-test-reg", 2, 70)]:
        arch = synthetic.VIT_ARCHS[arch_name]
        sd = synthetic.make_vit_state_dict(arch, seed=11)
        ext = build_tiny_reference(arch, sd, layer)
        images = synthetic.make_crops(2, (size, size), seed=12)
        with torch.no_grad():
            out = ext(images)
        gold[f"vit/{arch_name}/feature_maps"] = out["feature_maps"].contiguous().clone()
        gold[f"vit/{arch_name}/cls_tokens"] = out["cls_tokens"].contiguous().clone()
        gold[f"vit/{arch_name}/meta"] = {"layer": layer, "size": size, "wseed": 11, "iseed": 12}
        for facet in ("key", "value"):
            ext.facet = facet
            with torch.no_grad():
                outf = ext(images)
            gold[f"vit/{arch_name}/feature_maps_{facet}"] = outf["feature_maps"].contiguous().clone()
        ext.facet = "token"

    for model_name, tag in [("dinov2_version=vits14-reg_stride=14_facet=token_layer=9_logbin=0_norm=1", "vits14-reg"),
                            ("dinov2_vitl14", "vitl14")]:
        from oracle import vit as ovit

        opts = ovit.parse_extractor_name(model_name)
        arch = synthetic.VIT_ARCHS[opts["version"]]
        sd = synthetic.make_vit_state_dict(arch, seed=21)
        ext = build_reference_extractor(model_name, sd)
        images = synthetic.make_crops(1, (420, 420), seed=22)
        with torch.no_grad():
            out = ext(images)
        fm = out["feature_maps"]  # 1 x D x 30 x 30
        gold[f"vit/{tag}/feature_maps_sub"] = fm[:, ::8, ::3, ::3].contiguous().clone()
        gold[f"vit/{tag}/feature_maps_sum"] = fm.double().sum(dim=(2, 3)).float().clone()
        gold[f"vit/{tag}/cls_tokens"] = out["cls_tokens"].contiguous().clone()
        gold[f"vit/{tag}/meta"] = {"layer": ext.layer, "name": model_name, "wseed": 21, "iseed": 22}

    # ---- A5-A7: grid, mask filter, sampling ---------------------------------------------------
    grid14 = feature_util.generate_grid_points((420, 420), 14.0)
    gold["feature/grid14"] = grid14.clone()
    gold["feature/grid_56_4"] = feature_util.generate_grid_points((56, 40), 4.0).clone()
    mask = synthetic.make_masks(1, (420, 420), seed=31)[0]
    qp = feature_util.filter_points_by_mask(grid14, mask)
    gold["feature/mask_seed"] = 31
    gold["feature/filtered14"] = qp.clone()
    fmap = torch.randn(48, 30, 30, generator=torch.Generator().manual_seed(32))
    gold["feature/fmap_seed"] = 32
    gold["feature/sampled14"] = feature_util.sample_feature_map_at_points(fmap, qp, (420, 420)).contiguous().clone()
    pts = torch.rand(200, 2, generator=torch.Generator().manual_seed(33)) * 430.0 - 5.0
    gold["feature/random_points"] = pts.clone()
    gold["feature/sampled_random"] = feature_util.sample_feature_map_at_points(fmap, pts, (420, 420)).contiguous().clone()

    # ---- A8: PCA projection ---------------------------------------------------------------------
    pdict = synthetic.make_pca(128, 64, seed=41)
    projector = projector_util.projector_from_tensordict(pdict)
    x = synthetic.fp16_representable(torch.randn(300, 128, generator=torch.Generator().manual_seed(42)))
    gold["pca/in_seed"] = 42
    gold["pca/out"] = projector_util.project_features(x, [projector]).contiguous().clone()
    back = projector_util.projector_to_tensordict(projector)
    gold["pca/roundtrip_components_equal"] = bool(torch.equal(back["pca_projector"]["components"],
                                                              pdict["pca_projector"]["components"]))

    # ---- A9-A14: k-NN, tf-idf retrieval, cyclic correspondences -------------------------------
    bank = synthetic.make_bank_tensors(num_templates=24, patches_per_template=48, feat_dim=64,
                                       num_words=32, seed=51, ragged=True)
    feat = bank["feat_vectors"]
    centroids = bank["feat_cluster_centroids"]
    wk = knn_util.KNN(k=1, metric="l2")
    wk.fit(centroids)
    feat_to_word = wk.search(feat)[1].flatten()
    opts = repre_util.TemplateDescOpts()
    descs, idfs = template_util.calc_tfidf_descriptors(
        feat_vectors=feat, feat_to_word_ids=feat_to_word, feat_to_template_ids=bank["feat_to_template_ids"],
        feat_words=centroids, num_templates=24, tfidf_knn_k=opts.tfidf_knn_k,
        tfidf_soft_assign=opts.tfidf_soft_assign, tfidf_soft_sigma_squared=opts.tfidf_soft_sigma_squared)
    gold["bank/seed"] = 51
    gold["bank/feat_to_word"] = feat_to_word.clone()
    gold["bank/template_descs"] = descs.clone()
    gold["bank/idfs"] = idfs.clone()

    repre = repre_util.FeatureBasedObjectRepre(
        vertices=bank["vertices"], feat_vectors=feat, feat_to_template_ids=bank["feat_to_template_ids"],
        feat_to_vertex_ids=bank["feat_to_vertex_ids"], feat_cluster_centroids=centroids,
        feat_cluster_idfs=idfs, template_descs=descs, template_desc_opts=opts,
        feat_opts=repre_util.FeatureOpts(extractor_name="dinov2_vits14-reg"))

    # KNN class semantics.
    q = synthetic.make_query_features(150, 64, feat, seed=52)
    gold["knn/query_seed"] = 52
    k3 = knn_util.KNN(k=3, metric="l2")
    k3.fit(centroids)
    d3, i3 = k3.search(q)
    gold["knn/l2_k3_d"], gold["knn/l2_k3_i"] = d3.clone(), i3.clone()
    kc = knn_util.KNN(k=2, metric="cosine")
    kc.fit(feat)
    dc, ic = kc.search(q)
    gold["knn/cos_k2_d"], gold["knn/cos_k2_i"] = dc.clone(), ic.clone()

    # calc_tfidf hard / soft.
    wid, wd = template_util.find_nearest_object_features(q, k3)
    gold["tfidf/word_ids"], gold["tfidf/word_dists"] = wid.clone(), wd.clone()
    gold["tfidf/hard"] = template_util.calc_tfidf(wid, wd, idfs, soft_assignment=False).clone()
    gold["tfidf/soft"] = template_util.calc_tfidf(wid, wd, idfs, soft_assignment=True,
                                                  soft_sigma_squared=10.0).clone()
    tids, tscores = template_util.tfidf_matching(q, repre, 5, k3)
    gold["tfidf/top5_ids"], gold["tfidf/top5_scores"] = tids.clone(), tscores.clone()

    # establish_correspondences end to end (reference code, stub faiss).
    grid = feature_util.generate_grid_points((210, 140), 14.0)  # 15 x 10 = 150 query points
    template_knn = []
    for t in range(24):
        ids = torch.nonzero(bank["feat_to_template_ids"] == t).flatten()
        kt = knn_util.KNN(k=1, metric="l2")
        kt.fit(feat[ids])
        template_knn.append(kt)
    corresp = corresp_util.establish_correspondences(
        query_points=grid, query_features=q, object_repre=repre, template_matching_type="tfidf",
        feat_matching_type="cyclic_buddies", top_n_templates=5, top_k_buddies=40,
        visual_words_knn_index=k3, template_knn_indices=template_knn, debug=True)
    gold["corresp/grid"] = grid.clone()
    gold["corresp/list"] = [{k: (v.clone() if torch.is_tensor(v) else v) for k, v in c.items()} for c in corresp]

    out_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.pt")
    torch.save(gold, out_path)
    print(f"wrote {out_path}: {os.path.getsize(out_path) / 1e6:.2f} MB, {len(gold)} entries")


if __name__ == "__main__":
    main()
