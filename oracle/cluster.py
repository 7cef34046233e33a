"""Oracle: k-means as the reference obtains it from faiss.Kmeans.  TEST INFRASTRUCTURE ONLY.

Reference call site: utils/cluster_util.py:13-68 - `faiss.Kmeans(d, k, niter=50, seed=0, spherical=False)`,
`.train(samples)`, then `kmeans.index.search(samples, 1)`; used by scripts/gen_repre.py:289-300 to turn
the template features into 2048 visual words.  faiss 1.8.0 (conda_foundpose_gpu.yaml:19) is a third-party
dependency absent from /root/reference and from this image - **parity unpinned against faiss binaries**.
This file restates its published algorithm.  Step -> faiss 1.8.0 function it follows (faiss is not available
offline, so functions are named instead of line numbers that could not be verified here):

  rand_perm                   faiss/utils/random.cpp  `rand_perm(int* perm, size_t n, int64_t seed)`
  subsampling in `kmeans`     faiss/Clustering.cpp    `subsample_training_set` (called from `Clustering::train_encoded`
                                                      when nx > k * max_points_per_centroid, seed = cp.seed)
  initial centroids           faiss/Clustering.cpp    `Clustering::train_encoded`: `rand_perm(perm, nx, seed + 1 + redo)`,
                                                      centroid i = sample perm[i]  (nredo = 1 -> redo = 0)
  assign_nearest              faiss/Clustering.cpp    `index.search(nx, x, 1, dis, assign)` on an IndexFlatL2 holding
                                                      the centroids (exhaustive_L2sqr_blas, see oracle/knn.py)
  update_centroids            faiss/Clustering.cpp    `compute_centroids`: per-centroid sum of its samples / count
  split_clusters              faiss/Clustering.cpp    `split_clusters`: RandomGenerator rng(1234), EPS = 1/1024
  final assignment            faiss/Clustering.cpp    kmeans.index.search(samples, 1) at utils/cluster_util.py:54-56

Algorithm:

  * more than k*256 samples: train on the first k*256 entries of `rand_perm(n, seed)`;
  * initial centroids: the first k entries of `rand_perm(nx, seed + 1)`;
  * rand_perm: Fisher-Yates `swap(perm[i], perm[i + mt() % (n - i)])` with std::mt19937(seed)
    (numpy's legacy RandomState(seed) produces the same 32-bit stream);
  * 50 x { assign every sample to its nearest centroid (L2); centroid = mean of its samples;
           split_clusters: every empty cluster takes a copy of a cluster drawn with probability
           proportional to size (RandomGenerator(1234)), the two copies are perturbed by 1 +- 1/1024
           in alternating dimensions and share the size };
  * final assignment of ALL samples to the trained centroids.

Two choices make the restatement bit-reproducible on a GPU and are shared with the product
(foundpose_b200/utils/cluster_util.py): distances use fp16-rounded operands with fp32 arithmetic
(the storage format of the k-NN kernel; ties -> lower index), and cluster sums are accumulated in
64-bit fixed point (value * 2^24, round-half-even), which is order-independent.
"""

from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch

from oracle import knn as oknn

MAX_POINTS_PER_CENTROID = 256
EPS = np.float32(1.0 / 1024.0)
FIXED_SCALE = 16777216.0


def rand_perm(n: int, seed: int) -> np.ndarray:
    """faiss rand_perm (utils/random.cpp): mt19937(seed), perm[i] <-> perm[i + mt() % (n - i)]."""
    rs = np.random.RandomState(seed & 0xFFFFFFFF)
    raw = rs._bit_generator.random_raw(n)
    perm = np.arange(n, dtype=np.int64)
    steps = (raw % np.arange(n, 0, -1, dtype=np.uint64)).astype(np.int64)
    p = perm.tolist()
    st = steps.tolist()
    for i in range(n - 1):
        j = i + st[i]
        p[i], p[j] = p[j], p[i]
    return np.asarray(p, dtype=np.int64)


def assign_nearest(x: torch.Tensor, centroids: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(squared distance, id) of the nearest centroid; operands rounded to fp16, arithmetic fp32."""
    d, i = oknn.knn_l2(x.half().float(), centroids.half().float(), 1)
    return d[:, 0], i[:, 0]


def update_centroids(x: np.ndarray, assign: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """Mean per cluster through 64-bit fixed-point sums (bit-equal to fp_kmeans_update)."""
    fixed = np.rint(x.astype(np.float64) * FIXED_SCALE).astype(np.int64)
    sums = np.zeros((k, x.shape[1]), dtype=np.int64)
    np.add.at(sums, assign, fixed)
    counts = np.bincount(assign, minlength=k).astype(np.int32)
    cent = np.zeros((k, x.shape[1]), dtype=np.float32)
    nz = counts > 0
    cent[nz] = (sums[nz].astype(np.float64) / (FIXED_SCALE * counts[nz, None].astype(np.float64))).astype(np.float32)
    return cent, counts


def split_clusters(centroids: np.ndarray, hassign: np.ndarray, n: int) -> int:
    """faiss split_clusters (Clustering.cpp), in place.  Returns the number of splits."""
    k, d = centroids.shape
    rs = np.random.RandomState(1234)
    hassign = hassign.astype(np.float32)
    nsplit = 0
    for ci in range(k):
        if hassign[ci] != 0:
            continue
        cj = 0
        while True:
            p = np.float32(hassign[cj] - np.float32(1.0)) / np.float32(n - k)
            r = np.float32(rs._bit_generator.random_raw()) / np.float32(4294967295.0)
            if r < p:
                break
            cj = (cj + 1) % k
        centroids[ci] = centroids[cj]
        even = np.arange(d) % 2 == 0
        centroids[ci, even] *= np.float32(1) + EPS
        centroids[cj, even] *= np.float32(1) - EPS
        centroids[ci, ~even] *= np.float32(1) - EPS
        centroids[cj, ~even] *= np.float32(1) + EPS
        hassign[ci] = np.float32(int(hassign[cj]) // 2)
        hassign[cj] -= hassign[ci]
        nsplit += 1
    return nsplit


def kmeans(samples: torch.Tensor, num_centroids: int, num_iter: int = 50, seed: int = 0
           ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, List[float]]:
    """(centroids [k,d] fp32, cluster_ids [n] int32, squared centroid distances [n] fp32, objective per iteration)."""
    x = samples.detach().cpu().to(torch.float32).contiguous()
    n, d = x.shape
    k = num_centroids
    if n < k:
        raise ValueError(f"Number of training points ({n}) should be at least as large as number of clusters ({k})")
    xt = x
    if n > k * MAX_POINTS_PER_CENTROID:
        xt = x[torch.from_numpy(rand_perm(n, seed)[: k * MAX_POINTS_PER_CENTROID])]
    nx = xt.shape[0]
    centroids = xt[torch.from_numpy(rand_perm(nx, seed + 1)[:k])].clone()
    objective: List[float] = []
    xt_np = xt.numpy()
    for _ in range(num_iter):
        dist, ids = assign_nearest(xt, centroids)
        objective.append(float(dist.sum()))
        cent, counts = update_centroids(xt_np, ids.numpy(), k)
        split_clusters(cent, counts, nx)
        centroids = torch.from_numpy(cent)
    dist, ids = assign_nearest(x, centroids)
    return centroids, ids.to(torch.int32), dist, objective
