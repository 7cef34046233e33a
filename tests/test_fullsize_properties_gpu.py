"""Size-independent properties at BASELINE-like sizes (the oracle would take minutes here):
self-retrieval, sortedness, recomputed distances, idempotence of the k-NN, and template self-retrieval
through the whole retrieval + cyclic-matching stage."""
import pytest
import torch

from foundpose_b200 import synthetic

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big_bank():
    # 500 templates x 1024 patches x 256-d: 512k bank rows (config-2 row width and template size).
    g = torch.Generator(device="cuda").manual_seed(0)
    T, P, d, W = 500, 1024, 256, 2048
    feat = torch.randn(T * P, d, device="cuda", generator=g).half().float()
    tpl = torch.arange(T, dtype=torch.int32, device="cuda").repeat_interleave(P)
    verts = torch.randn(T * P, 3, device="cuda", generator=g)
    cent = feat[torch.randperm(T * P, device="cuda", generator=g)[:W]].clone()
    return dict(T=T, P=P, d=d, W=W, feat=feat, tpl=tpl, verts=verts, cent=cent)


def test_knn_fullsize_properties(big_bank):
    from foundpose_b200.utils import knn_util

    feat = big_bank["feat"]
    index = knn_util.KNN(k=5, metric="l2")
    index.fit(feat)
    ids = torch.randperm(feat.shape[0], device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))[:3000]
    q = feat[ids]
    d, i = index.search(q)
    # a bank row retrieves itself at distance ~0, results are sorted ascending
    assert torch.equal(i[:, 0], ids)
    assert float(d[:, 0].abs().max()) <= 1e-2          # fp32 cancellation of ||q||^2 + ||x||^2 - 2<q,x> at norm ~256
    assert bool((d[:, 1:] >= d[:, :-1]).all())
    # every returned distance equals the recomputed squared distance to the returned row (1e-3 relative)
    rec = (q.unsqueeze(1) - feat[i]).square().sum(-1)
    assert bool(((d - rec).abs() <= 1e-3 * rec + 1e-2).all())
    # split-bank (few queries) and dense (many queries) paths agree
    d2, i2 = index.search(q[:100])
    assert torch.equal(i2, i[:100]) and torch.allclose(d2, d[:100], rtol=1e-4, atol=1e-3)
    # idempotence
    d3, i3 = index.search(q)
    assert torch.equal(i3, i) and torch.equal(d3, d)


def test_template_self_retrieval_and_cyclic_identity(big_bank):
    """A crop whose descriptors ARE template t's descriptors must retrieve t first with score ~1, and the
    cyclic matching against t is the identity: all cycle distances 0, 3D points = t's vertices."""
    from foundpose_b200 import _native, pipeline
    from foundpose_b200.utils import knn_util, repre_util, template_util

    b = big_bank
    wk = knn_util.KNN(1, "l2"); wk.fit(b["cent"])
    f2w = wk.search(b["feat"])[1].flatten()
    descs, idfs = template_util.calc_tfidf_descriptors(b["feat"], f2w, b["tpl"], b["cent"], b["T"], 3, False, 10.0)
    assert torch.isfinite(descs).all()
    repre = repre_util.FeatureBasedObjectRepre(
        vertices=b["verts"], feat_vectors=b["feat"], feat_to_template_ids=b["tpl"], feat_cluster_centroids=b["cent"],
        feat_cluster_idfs=idfs, template_descs=descs, template_desc_opts=repre_util.TemplateDescOpts())
    index = pipeline.ObjectIndex(repre, torch.device("cuda"))
    B, stride, nq = 8, 900, 900
    engine = pipeline.RetrievalEngine(index, B, stride, 5, 300)
    templates = [3, 77, 123, 250, 311, 404, 480, 499]
    feats = torch.zeros(B * stride, b["d"], device="cuda")
    pts = torch.zeros(B, stride, 2, device="cuda")
    grid = torch.stack(torch.meshgrid(torch.arange(30.), torch.arange(30.), indexing="xy"), -1).reshape(-1, 2) * 14 + 7
    for k, t in enumerate(templates):
        feats[k * stride: k * stride + nq] = b["feat"][t * b["P"]: t * b["P"] + nq]
        pts[k, :nq] = grid.cuda()
    cnt = torch.full((B,), nq, dtype=torch.int32, device="cuda")
    out = engine.match(_native.convert_rows_f16(feats), pts, cnt)
    torch.cuda.synchronize()
    assert out.template_ids[:, 0].tolist() == templates
    assert float(out.template_scores[:, 0].min()) > 0.85      # 900 of the template's 1024 patches
    assert bool((out.template_scores[:, :-1] >= out.template_scores[:, 1:]).all())
    assert out.count[:, 0].tolist() == [300] * B
    assert float(out.dists[:, 0].abs().max()) == 0.0          # cycle closes on itself
    for k, t in enumerate(templates):
        qid = out.query_ids[k, 0]
        assert torch.equal(qid, torch.arange(300, device="cuda"))         # ties -> ascending query id
        assert torch.equal(out.vertex_ids[k, 0], t * b["P"] + qid)
        assert torch.equal(out.coord_3d[k, 0], b["verts"][t * b["P"] + qid])
        assert torch.equal(out.coord_2d[k, 0], pts[k, :300])
