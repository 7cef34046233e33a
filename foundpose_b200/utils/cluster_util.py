"""K-means on the GPU - drop-in for the reference's utils/cluster_util.py (faiss.Kmeans wrapper).

`kmeans(samples, num_centroids, num_iter=50, verbose=True) -> (centroids, cluster_ids, centroid_distances)`
with the reference's return types (utils/cluster_util.py:13-68: fp32 centroids, int32 ids, fp32 squared
distances, all on the device of `samples`).  Used by the offline bank build (scripts/gen_repre.py:289-300)
to quantise the template features into visual words.

faiss's training loop (Clustering.cpp, seed 0, niter 50, at most 256 samples per centroid, empty clusters
split) is restated here; see oracle/cluster.py for the list of what is kept.  The two heavy steps run on
the hand-written kernels: the assignment is the visual-word search K1 (`fp_knn_search_items`, tcgen05
distance tiles over fp16 rows) and the update is `fp_kmeans_update` (order-independent fixed-point
sums).  The random permutations and the rare empty-cluster splits are host logic on k integers.
"""

from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from foundpose_b200 import _native
from foundpose_b200.utils import knn_util, logging

logger: logging.Logger = logging.get_logger()

MAX_POINTS_PER_CENTROID = 256      # faiss ClusteringParameters default
_EPS = np.float32(1.0 / 1024.0)


def rand_perm(n: int, seed: int) -> np.ndarray:
    """faiss rand_perm (utils/random.cpp): std::mt19937(seed), perm[i] <-> perm[i + mt() % (n - i)]."""
    rs = np.random.RandomState(seed & 0xFFFFFFFF)      # legacy seeding == std::mt19937(seed)
    raw = rs._bit_generator.random_raw(n)
    steps = (raw % np.arange(n, 0, -1, dtype=np.uint64)).astype(np.int64).tolist()
    perm = list(range(n))
    for i in range(n - 1):
        j = i + steps[i]
        perm[i], perm[j] = perm[j], perm[i]
    return np.asarray(perm, dtype=np.int64)


def _split_clusters(centroids: np.ndarray, hassign: np.ndarray, n: int) -> int:
    """faiss split_clusters (Clustering.cpp) on host copies; returns the number of splits."""
    k, d = centroids.shape
    rs = np.random.RandomState(1234)
    hassign = hassign.astype(np.float32)
    even = np.arange(d) % 2 == 0
    nsplit = 0
    for ci in np.nonzero(hassign == 0)[0].tolist():
        cj = 0
        while True:
            p = np.float32(hassign[cj] - np.float32(1.0)) / np.float32(n - k)
            r = np.float32(rs._bit_generator.random_raw()) / np.float32(4294967295.0)
            if r < p:
                break
            cj = (cj + 1) % k
        centroids[ci] = centroids[cj]
        centroids[ci, even] *= np.float32(1) + _EPS
        centroids[cj, even] *= np.float32(1) - _EPS
        centroids[ci, ~even] *= np.float32(1) - _EPS
        centroids[cj, ~even] *= np.float32(1) + _EPS
        hassign[ci] = np.float32(int(hassign[cj]) // 2)
        hassign[cj] -= hassign[ci]
        nsplit += 1
    return nsplit


def kmeans(samples: torch.Tensor, num_centroids: int, num_iter: int = 50, verbose: bool = True,
           seed: int = 0) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """K-means clustering; returns (centroids, cluster_ids, centroid_distances) like the reference."""
    if not samples.is_cuda:
        raise ValueError("foundpose_b200 runs on CUDA tensors only (no CPU fallback)")
    dev = samples.device
    x = samples.detach().to(torch.float32).contiguous()
    n, d = x.shape
    k = int(num_centroids)
    if n < k:
        raise ValueError(f"Number of training points ({n}) should be at least as large as number of clusters ({k})")
    xt = x
    if n > k * MAX_POINTS_PER_CENTROID:
        if verbose:
            logger.info(f"Sampling a subset of {k * MAX_POINTS_PER_CENTROID} / {n} for training")
        sel = torch.from_numpy(rand_perm(n, seed)[: k * MAX_POINTS_PER_CENTROID]).to(dev)
        xt = x[sel].contiguous()
    nx = xt.shape[0]
    centroids = xt[torch.from_numpy(rand_perm(nx, seed + 1)[:k]).to(dev)].contiguous()
    sums = torch.empty(k * d, dtype=torch.int64, device=dev)
    counts = torch.empty(k, dtype=torch.int32, device=dev)
    index = knn_util.KNN(k=1, metric="l2")
    for it in range(num_iter):
        index.fit(centroids)
        dist, ids = index.search(xt)
        _native.kmeans_update(xt, ids.reshape(-1).contiguous(), k, sums, counts, centroids)
        hassign = counts.cpu().numpy()
        nsplit = 0
        if (hassign == 0).any():
            cent = centroids.cpu().numpy()
            nsplit = _split_clusters(cent, hassign, nx)
            centroids.copy_(torch.from_numpy(cent))
        if verbose:
            logger.info(f"  Iteration {it} objective={float(dist.sum()):.6g} split={nsplit}")
    index.fit(centroids)
    dist, ids = index.search(x)
    return centroids, ids.reshape(-1).to(torch.int32), dist.reshape(-1)
