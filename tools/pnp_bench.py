#!/usr/bin/env python3
"""Coarse-pose step timing: fp_pnp_ransac on the GPU vs the reference's OpenCV calls on the host cores.

Workload = BASELINE configs[1] shape: 64 crops x 5 templates = 320 problems x 300 correspondences
(50% gross outliers, 0.3 px noise), 400 RANSAC iterations (configs/infer/lmo.json) and 1000 (InferOpts default).
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from foundpose_b200.utils import pnp_util  # noqa: E402

P, M = 320, 300


def rodrigues(w):
    th = float(np.linalg.norm(w))
    K = np.array([[0.0, -w[2], w[1]], [w[2], 0.0, -w[0]], [-w[1], w[0], 0.0]])
    return np.eye(3) + np.sin(th) / th * K + (1.0 - np.cos(th)) / (th * th) * (K @ K)


rng = np.random.default_rng(11)
X = (rng.normal(size=(P, M, 3)) * 60).astype(np.float32)
K4 = np.tile(np.array([600.0, 600.0, 210.0, 210.0]), (P, 1))
Rs = np.stack([rodrigues(rng.normal(size=3)) for _ in range(P)])
ts = np.stack([np.array([rng.uniform(-40, 40), rng.uniform(-40, 40), rng.uniform(500, 900)]) for _ in range(P)])
Xc = np.einsum("pij,pmj->pmi", Rs, X.astype(np.float64)) + ts[:, None, :]
x = np.stack([600.0 * Xc[..., 0] / Xc[..., 2] + 210.0, 600.0 * Xc[..., 1] / Xc[..., 2] + 210.0], -1)
x += rng.normal(size=x.shape) * 0.3
out = rng.random((P, M)) < 0.5
x[out] += rng.choice([-1.0, 1.0], size=(int(out.sum()), 2)) * rng.uniform(40, 200, size=(int(out.sum()), 2))
x = x.astype(np.float32)

dev = torch.device("cuda", 0)
c2d, c3d = torch.from_numpy(x).to(dev), torch.from_numpy(X).to(dev)
cnt = torch.full((P,), M, dtype=torch.int32, device=dev)
Kd = torch.from_numpy(K4).to(dev)
res = {"problems": P, "correspondences": M}
for iters in (400, 1000):
    for conf, tag in ((0.99, "conf0.99"), (1.0, "all_iters")):
        for _ in range(3):
            pnp_util.estimate_poses_batched(c2d, c3d, cnt, Kd, iters, 10.0, conf)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            r = pnp_util.estimate_poses_batched(c2d, c3d, cnt, Kd, iters, 10.0, conf)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        res[f"gpu_ms_iters{iters}_{tag}"] = round(ms, 4)
        res[f"gpu_problems_per_s_iters{iters}_{tag}"] = round(P / ms * 1e3, 1)
        assert bool(r["success"].all())
try:
    import cv2
    cv2.setNumThreads(os.cpu_count() or 1)
    K = np.array([[600.0, 0, 210.0], [0, 600.0, 210.0], [0, 0, 1]])
    n = 64
    t0 = time.perf_counter()
    for p in range(n):
        ok, rv, tv, inl = cv2.solvePnPRansac(X[p], x[p], K, None, iterationsCount=400, reprojectionError=10.0,
                                             confidence=0.99, flags=cv2.SOLVEPNP_ITERATIVE)
        cv2.solvePnPRefineLM(X[p][inl[:, 0]], x[p][inl[:, 0]], K, None, rv, tv)
    dt = time.perf_counter() - t0
    res["opencv_cpu_ms_per_problem_iters400"] = round(dt / n * 1e3, 4)
    res["opencv_cpu_problems_per_s"] = round(n / dt, 1)
    res["cpu_cores"] = os.cpu_count()
except ImportError:
    res["opencv"] = "not installed on this box"
print(json.dumps(res))
