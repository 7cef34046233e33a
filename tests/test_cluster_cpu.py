"""oracle/cluster.py and the host logic of foundpose_b200/utils/cluster_util.py (no GPU)."""
import numpy as np
import torch

from oracle import cluster as ocluster


def test_rand_perm_is_mt19937_fisher_yates():
    # std::mt19937(0) starts 2357136044, 2546248239, 3071714933 (C++ standard engine, seed 0)
    raw = np.random.RandomState(0)._bit_generator.random_raw(3)
    assert raw.tolist() == [2357136044, 2546248239, 3071714933]
    p = ocluster.rand_perm(10, 0)
    assert sorted(p.tolist()) == list(range(10))
    assert p[0] == 0 + 2357136044 % 10 and np.array_equal(p, ocluster.rand_perm(10, 0))
    from foundpose_b200.utils import cluster_util
    assert np.array_equal(cluster_util.rand_perm(1000, 5), ocluster.rand_perm(1000, 5))   # host logic == oracle


def test_update_and_split():
    rng = np.random.default_rng(0)
    x = rng.normal(size=(200, 8)).astype(np.float32)
    assign = rng.integers(0, 5, size=200)
    assign[assign == 3] = 2                       # cluster 3 empty
    cent, counts = ocluster.update_centroids(x, assign, 5)
    assert counts[3] == 0 and not cent[3].any()
    for c in (0, 1, 2, 4):
        assert np.allclose(cent[c], x[assign == c].mean(0), atol=1e-6)
    before = cent.copy()
    assert ocluster.split_clusters(cent, counts, 200) == 1
    donor = [c for c in (0, 1, 2, 4) if not np.array_equal(before[c], cent[c])]
    assert len(donor) == 1
    assert np.allclose(cent[3], before[donor[0]], rtol=2e-3) and not np.array_equal(cent[3], cent[donor[0]])


def test_kmeans_recovers_separated_blobs():
    g = torch.Generator().manual_seed(0)
    centers = torch.randn(6, 16, generator=g) * 10
    x = (centers[torch.arange(600) % 6] + 0.1 * torch.randn(600, 16, generator=g))
    cent, ids, dist, obj = ocluster.kmeans(x, 6, num_iter=10)
    assert ids.dtype == torch.int32 and cent.shape == (6, 16) and dist.shape == (600,)
    assert all(b <= a * (1 + 1e-6) for a, b in zip(obj[1:], obj[2:]))        # Lloyd iterations never increase the objective
    # no blob is split over two clusters unless a cluster sits between blobs: every cluster is pure
    for c in range(6):
        members = (torch.arange(600) % 6)[ids == c]
        assert members.numel() == 0 or len(set(members.tolist())) <= 2
    try:
        ocluster.kmeans(x[:3], 6)
        assert False
    except ValueError:
        pass
