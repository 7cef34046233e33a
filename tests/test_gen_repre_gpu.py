"""Offline bank build on the GPU (foundpose_b200/scripts/gen_repre.py) end to end on a tiny extractor:
raw features registered in 3D against the oracle, the repre.pth round trip, and self-retrieval of the templates."""
import numpy as np
import pytest
import torch

from foundpose_b200 import synthetic

pytestmark = pytest.mark.gpu

ARCH = "tiny-test-reg"
LAYER = 2
SIZE = 98


def _templates(num: int, seed: int = 0):
    from foundpose_b200.scripts import gen_repre
    from foundpose_b200.utils import structs
    from oracle import pnp as opnp
    rng = np.random.default_rng(seed)
    images = synthetic.make_crops(num, (SIZE, SIZE), seed=seed + 1)
    yy, xx = torch.meshgrid(torch.arange(SIZE), torch.arange(SIZE), indexing="ij")
    out = []
    for t in range(num):
        cx, cy, r = rng.uniform(35, 63), rng.uniform(35, 63), rng.uniform(25, 40)
        mask = (((xx - cx) ** 2 + (yy - cy) ** 2) < r * r).to(torch.float32)
        depth = 400.0 + 30.0 * torch.sin(xx / 9.0 + t) + 20.0 * torch.cos(yy / 7.0)
        R = opnp.rodrigues(rng.normal(size=3))
        T_wc = np.eye(4)
        T_wc[:3, :3], T_wc[:3, 3] = R, rng.normal(size=3) * 50
        cam = structs.PinholePlaneCameraModel(SIZE, SIZE, (110.0, 105.0), (48.0, 50.0), T_world_from_eye=T_wc)
        T_wm = np.eye(4)
        T_wm[:3, :3], T_wm[:3, 3] = opnp.rodrigues(rng.normal(size=3)), rng.normal(size=3) * 20
        out.append(gen_repre.TemplateSample(images[t], depth, mask, cam, T_wm))
    return out


def _opts():
    from foundpose_b200.scripts import gen_repre
    from foundpose_b200.utils import repre_util
    return gen_repre.GenRepreOpts(
        extractor_name=f"dinov2_version={ARCH}_stride=14_facet=token_layer={LAYER}_norm=1", grid_cell_size=14.0,
        pca_components=64, cluster_num=16, template_desc_opts=repre_util.TemplateDescOpts(), debug=False)


def _extractor():
    from foundpose_b200.utils import dinov2_utils
    arch = synthetic.VIT_ARCHS[ARCH]
    sd = synthetic.make_vit_state_dict(arch, seed=3)
    return dinov2_utils.DinoFeatureExtractor(_opts().extractor_name, state_dict=sd, max_batch=32), arch, sd


def test_raw_repre_matches_oracle():
    from foundpose_b200.scripts import gen_repre
    from oracle import feature as ofeature
    from oracle import vit as ovit
    dev = torch.device("cuda", 0)
    templates = _templates(5)
    ext, arch, sd = _extractor()
    ext.to(dev)
    repre = gen_repre.generate_raw_repre(_opts(), templates, ext, dev, extract_batch=4)
    tpl = repre.feat_to_template_ids.cpu()
    assert tpl.dtype == torch.int32 and bool((tpl[1:] >= tpl[:-1]).all())
    assert repre.templates.shape == (5, 3, SIZE, SIZE) and repre.templates.dtype == torch.uint8
    grid = ofeature.generate_grid_points((SIZE, SIZE), 14.0)
    for t, s in enumerate(templates):
        # oracle: 5x5 erosion = every pixel of the window set (window clipped at the border)
        m = s.object_mask
        er = -torch.nn.functional.max_pool2d(-m[None, None], 5, 1, 2)[0, 0]
        pts = ofeature.filter_points_by_mask(grid, er)
        fmap = ovit.extract(sd, arch, s.image_chw[None], layer=LAYER)["feature_maps"][0]
        feat = ofeature.sample_feature_map_at_points(fmap, pts, (SIZE, SIZE))
        mine = repre.feat_vectors[(tpl == t).to(dev)].cpu()
        assert mine.shape == feat.shape
        assert float((mine - feat).norm() / feat.norm()) <= 5e-3
        # 3D registration in float64 numpy
        p = pts.numpy().astype(np.float64)
        focal = 0.5 * (s.camera.f[0] + s.camera.f[1])
        d = s.depth_image_hw.numpy()[np.floor(p[:, 1]).astype(int), np.floor(p[:, 0]).astype(int)].astype(np.float64)
        cam_pts = np.concatenate([p - np.array(s.camera.c), np.full((len(p), 1), focal)], 1) * (d / focal)[:, None]
        T_mc = np.linalg.inv(s.T_world_from_model) @ s.camera.T_world_from_eye
        ref_v = cam_pts @ T_mc[:3, :3].T + T_mc[:3, 3]
        assert np.abs(repre.vertices[(tpl == t).to(dev)].cpu().numpy() - ref_v).max() < 2e-2     # fp32 vs fp64, |v| ~ 500


def test_generate_repre_round_trip_and_self_retrieval(tmp_path):
    from foundpose_b200.scripts import gen_repre
    from foundpose_b200.utils import corresp_util, feature_util, knn_util, projector_util, repre_util
    dev = torch.device("cuda", 0)
    templates = _templates(24)
    ext, _, _ = _extractor()
    opts = _opts()
    repre = gen_repre.generate_repre(opts, templates, device="cuda:0", extractor=ext, output_dir=str(tmp_path))
    F = repre.feat_vectors.shape[0]
    assert repre.feat_vectors.shape[1] == 64 and repre.feat_cluster_centroids.shape == (16, 64)
    assert repre.feat_to_cluster_ids.shape == (F,) and repre.feat_to_cluster_ids.dtype == torch.int32
    assert repre.template_descs.shape == (24, 16) and repre.feat_cluster_idfs.shape == (16,)
    assert len(repre.feat_raw_projectors) == 1 and repre.vertices.shape == (F, 3)
    # every feature sits in its nearest visual word
    index = knn_util.KNN(1, "l2")
    index.fit(repre.feat_cluster_centroids)
    assert torch.equal(index.search(repre.feat_vectors)[1][:, 0].to(torch.int32), repre.feat_to_cluster_ids)
    # repre.pth round trip through the reference's file layout
    loaded = repre_util.load_object_repre(str(tmp_path), tensor_device="cuda:0")
    for name in ("vertices", "feat_vectors", "feat_to_template_ids", "feat_cluster_centroids", "feat_cluster_idfs",
                 "template_descs", "feat_to_cluster_ids"):
        assert torch.equal(getattr(loaded, name).cpu(), getattr(repre, name).cpu()), name
    assert loaded.template_desc_opts == repre.template_desc_opts
    # the bank retrieves its own templates: query = a template's own image through the inference path
    hits = 0
    grid = feature_util.generate_grid_points((SIZE, SIZE), 14.0).to(dev)
    for t in range(0, 24, 3):
        s = templates[t]
        fmap = ext(s.image_chw[None].to(dev))["feature_maps"][0]
        pts = feature_util.filter_points_by_mask(grid, feature_util.erode_mask_5x5(s.object_mask.to(dev)))
        q = feature_util.sample_feature_map_at_points(fmap, pts, (SIZE, SIZE))
        q = projector_util.project_features(q, loaded.feat_raw_projectors).contiguous()
        corresp = corresp_util.establish_correspondences(pts, q, loaded, "tfidf", "cyclic_buddies", 3, 20)
        hits += int(int(corresp[0]["template_id"]) == t)
        if int(corresp[0]["template_id"]) == t:
            # cyclic buddies of a template with itself: every query maps to its own 3D point
            mine = loaded.vertices.cpu()[(loaded.feat_to_template_ids.cpu() == t)]
            assert torch.allclose(corresp[0]["coord_3d"].cpu(), mine[corresp[0]["coord_2d_ids"].cpu()], atol=1e-4)
    assert hits >= 7                                                        # 8 probes


@pytest.mark.parametrize("n,d,k", [(3000, 256, 64), (1111, 200, 16)])
def test_pca_fit_gpu_matches_sklearn(n, d, k):
    """PCAProjector.fit on CUDA data (tcgen05 covariance + eigh) against scikit-learn's exact solver."""
    from sklearn.decomposition import PCA

    from foundpose_b200.utils import projector_util
    g = torch.Generator().manual_seed(7)
    scales = torch.linspace(6.0, 1.0, 32)                       # 32 well separated directions + isotropic noise
    x = (torch.randn(n, 32, generator=g) * scales) @ torch.linalg.qr(torch.randn(d, 32, generator=g))[0].t()
    x = x + 0.05 * torch.randn(n, d, generator=g) + 3.0 * torch.randn(1, d, generator=g)
    proj = projector_util.PCAProjector(n_components=k)
    proj.fit(x.cuda())
    ref = PCA(n_components=k, svd_solver="full").fit(x.numpy())
    assert proj.pca.components_.shape == (k, d) and proj.pca.mean_.shape == (d,)
    assert np.allclose(proj.pca.mean_, ref.mean_, atol=1e-5)
    # eigenvalues agree to fp32 resolution of the LARGEST one (the covariance is accumulated in fp32): the 32 signal
    # eigenvalues to ~1e-5 relative, the noise floor (1e-4 of the top) to a fraction of a percent
    ev0 = float(ref.explained_variance_[0])
    assert np.allclose(proj.pca.explained_variance_, ref.explained_variance_, rtol=1e-4, atol=3e-6 * ev0)
    assert np.allclose(proj.pca.explained_variance_ratio_, ref.explained_variance_ratio_, rtol=1e-4,
                       atol=3e-6 * float(ref.explained_variance_ratio_[0]))
    assert np.allclose(proj.pca.singular_values_ ** 2 / (n - 1), proj.pca.explained_variance_, rtol=1e-5)
    assert np.allclose(proj.pca.singular_values_[:16], ref.singular_values_[:16], rtol=1e-4)
    assert abs(proj.pca.noise_variance_ - ref.noise_variance_) <= 2e-2 * ref.noise_variance_
    top = min(k, 32)                                            # the separated directions: same vectors, same signs
    dots = (proj.pca.components_[:top] * ref.components_[:top]).sum(1)
    assert dots.min() > 0.999
    # the projector is usable as fitted: transform == sklearn's transform on the separated directions
    if d % 64 == 0:       # the projection kernel takes feature widths that are multiples of 64 (all DINOv2 widths)
        mine = proj.transform(x.cuda()).cpu().numpy()[:, :top]
        theirs = ref.transform(x.numpy())[:, :top]
        assert np.abs(mine - theirs).max() <= 5e-3 * np.abs(theirs).max()
    # tensordict round trip (repre.pth layout)
    again = projector_util.projector_from_tensordict(projector_util.projector_to_tensordict(proj))
    assert np.array_equal(again.pca.components_, proj.pca.components_)
