"""GPU k-means (utils/cluster_util.py: fp_knn_search_items + fp_kmeans_update) against oracle/cluster.py."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_update_is_bit_exact_and_order_independent():
    from foundpose_b200 import _native
    from oracle import cluster as ocluster
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(3)
    n, d, k = 50000, 96, 37
    x = torch.randn(n, d, generator=g) * 3
    assign = torch.randint(0, k, (n,), generator=g)
    assign[assign == 5] = 6                      # an empty cluster
    assign[:7] = -1                              # unassigned rows are ignored
    sums = torch.empty(k * d, dtype=torch.int64, device=dev)
    counts = torch.empty(k, dtype=torch.int32, device=dev)
    cent = torch.empty(k, d, device=dev)
    _native.kmeans_update(x.to(dev), assign.to(dev), k, sums, counts, cent)
    ref_c, ref_n = ocluster.update_centroids(x[7:].numpy(), assign[7:].numpy(), k)
    assert np.array_equal(counts.cpu().numpy(), ref_n)
    assert np.array_equal(cent.cpu().numpy(), ref_c)                      # bit-exact, incl. the zero row
    perm = torch.randperm(n, generator=g)
    cent2 = torch.empty(k, d, device=dev)
    _native.kmeans_update(x[perm].to(dev), assign[perm].to(dev), k, sums, counts, cent2)
    assert torch.equal(cent, cent2)                                       # atomics order does not matter


def test_assignment_step_matches_oracle_outside_near_ties():
    from foundpose_b200.utils import knn_util
    from oracle import cluster as ocluster
    from oracle import knn as oknn
    g = torch.Generator().manual_seed(4)
    x = torch.randn(4000, 128, generator=g)
    cent = torch.randn(200, 128, generator=g) * 0.7
    index = knn_util.KNN(1, "l2")
    index.fit(cent.cuda())
    dist, ids = index.search(x.cuda())
    rd, ri = ocluster.assign_nearest(x, cent)
    margin = oknn.topk_margin(x.half().float(), cent.half().float(), 1)[:, 0]
    safe = margin > 1e-4
    assert safe.float().mean() > 0.99
    assert torch.equal(ids.cpu()[:, 0][safe], ri[safe])
    assert torch.allclose(dist.cpu()[:, 0], rd, rtol=1e-3, atol=1e-3)


def test_kmeans_separated_blobs_agrees_with_oracle_end_to_end():
    from foundpose_b200.utils import cluster_util
    from oracle import cluster as ocluster
    g = torch.Generator().manual_seed(0)
    centers = (torch.randn(8, 64, generator=g) * 10).half().float()
    x = (centers[torch.arange(2400) % 8] + 0.1 * torch.randn(2400, 64, generator=g))
    cent, ids, dist = cluster_util.kmeans(x.cuda(), 8, num_iter=10, verbose=False)
    rc, ri, rd, _ = ocluster.kmeans(x, 8, num_iter=10)
    assert ids.dtype == torch.int32 and ids.is_cuda and cent.is_cuda and dist.is_cuda
    # Same permutations, same update arithmetic.  |x|^2 ~ 6400 against within-blob distances ~ 1: the
    # ||q||^2 + ||x||^2 - 2<q,x> form (faiss's too) cancels to ~1e-3 absolute, so assignments inside a blob that
    # holds two centroids can differ between tensor-core and BLAS accumulation order and the runs drift apart
    # slightly; the clustering itself must be the same.
    assert (ids.cpu() == ri).float().mean() > 0.95
    assert abs(float(dist.sum()) - float(rd.sum())) < 2e-3 * float(rd.sum())
    assert torch.allclose(cent.cpu(), rc, atol=0.05)


def test_kmeans_subsampling_and_objective():
    from foundpose_b200.utils import cluster_util
    from oracle import cluster as ocluster
    g = torch.Generator().manual_seed(1)
    x = torch.randn(6000, 64, generator=g)            # 6000 > 16 * 256 -> trains on a 4096-sample subset
    cent, ids, dist = cluster_util.kmeans(x.cuda(), 16, num_iter=8, verbose=False)
    rc, ri, rd, _ = ocluster.kmeans(x, 16, num_iter=8)
    assert cent.shape == (16, 64) and ids.shape == (6000,) and int(ids.min()) >= 0 and int(ids.max()) < 16
    # unstructured data: near-ties can flip single assignments, the clustering quality must agree
    assert abs(float(dist.sum()) - float(rd.sum())) < 5e-3 * float(rd.sum())
    assert (ids.cpu() == ri).float().mean() > 0.97
    with pytest.raises(ValueError):
        cluster_util.kmeans(x[:5].cuda(), 16)
    with pytest.raises(ValueError):
        cluster_util.kmeans(x, 16)                     # CPU tensor: no fallback


def test_kmeans_is_exact_against_the_oracle_when_every_assignment_has_a_margin():
    """fp16-representable samples in well separated blobs, arranged so that faiss's initialisation (the first k
    entries of rand_perm(n, seed + 1)) picks one sample per blob: every assignment of every iteration then has a
    large margin, and the whole run must equal the oracle bit for bit - ids, centroids (fixed-point update) - with
    distances within 1e-3 relative."""
    from foundpose_b200.utils import cluster_util
    from oracle import cluster as ocluster

    k, n, d = 12, 3000, 64
    g = torch.Generator().manual_seed(5)
    centers = (torch.randn(k, d, generator=g) * 8).half().float()
    blob = torch.randint(0, k, (n,), generator=g)
    init = ocluster.rand_perm(n, 0 + 1)[:k]                   # n <= k * 256: no sub-sampling, seed 0
    blob[torch.from_numpy(init)] = torch.arange(k)            # one initial centroid per blob
    x = (centers[blob] + 0.25 * torch.randn(n, d, generator=g)).half().float()
    cent, ids, dist = cluster_util.kmeans(x.cuda(), k, num_iter=6, verbose=False)
    rc, ri, rd, _ = ocluster.kmeans(x, k, num_iter=6)
    assert torch.equal(ri.to(torch.int64), blob)              # the clustering is the blob structure
    assert torch.equal(ids.cpu(), ri)                         # bit-exact assignments
    assert np.array_equal(cent.cpu().numpy(), rc.numpy())     # bit-exact centroids
    # ||x||^2 ~ 4000 against distances ~ 2-5: the expansion ||q||^2 + ||x||^2 - 2<q,x> (faiss's too) cancels to ~1e-3
    # absolute in fp32, whatever the accumulation order
    assert torch.allclose(dist.cpu(), rd, rtol=1e-3, atol=4e-3)
