#!/usr/bin/env python3
"""Offline visual-word clustering timing: cluster_util.kmeans on the GPU at the LM-O scale the
reference trains on (faiss sub-samples to 256 points per centroid: 524,288 x 256-d samples, 2048 centroids, 50
iterations), next to ONE Lloyd iteration of the same arithmetic in torch on the host cores (bounded sample)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from foundpose_b200.utils import cluster_util  # noqa: E402

n, d, k, iters = 700000, 256, 2048, 50
g = torch.Generator().manual_seed(0)
x = torch.randn(n, d, generator=g)
xd = x.cuda()
cluster_util.kmeans(xd[:20000], 64, num_iter=2, verbose=False)      # warm-up (kernels, allocator)
torch.cuda.synchronize()
t0 = time.perf_counter()
cent, ids, dist = cluster_util.kmeans(xd, k, num_iter=iters, verbose=False)
torch.cuda.synchronize()
gpu_s = time.perf_counter() - t0
res = {"samples": n, "trained_on": k * 256, "dim": d, "centroids": k, "iterations": iters, "gpu_total_s": round(gpu_s, 3),
       "gpu_ms_per_iteration_incl_host_logic": round(gpu_s / (iters + 1) * 1e3, 2),
       "objective_per_sample": round(float(dist.mean()), 4), "empty_clusters": int((torch.bincount(ids.long(), minlength=k) == 0).sum())}
# host baseline: one assignment + update iteration on the training subset (what faiss does 50 times)
torch.set_num_threads(os.cpu_count() or 1)
xt = x[: k * 256]
c0 = xt[:k].clone()
t0 = time.perf_counter()
best = torch.empty(xt.shape[0], dtype=torch.int64)
for s in range(0, xt.shape[0], 8192):
    q = xt[s:s + 8192]
    dd = (q * q).sum(1, keepdim=True) + (c0 * c0).sum(1)[None] - 2.0 * (q @ c0.t())
    best[s:s + 8192] = dd.argmin(1)
sums = torch.zeros(k, d).index_add_(0, best, xt)
cpu_s = time.perf_counter() - t0
res.update({"cpu_s_per_iteration": round(cpu_s, 3), "cpu_cores": os.cpu_count(),
            "cpu_projected_total_s": round(cpu_s * (iters + 1), 1)})
print(json.dumps(res))
