#!/usr/bin/env python3
"""Golden vectors for the coarse-pose step, produced by OpenCV itself (the dependency that holds the
reference's arithmetic for utils/pnp_util.py:42-72).

Run in the build container (cv2 4.13):  python tests/golden/make_golden_pnp.py
Writes tests/golden/golden_pnp_v1.npz: seeded synthetic 2D-3D correspondences (known pose, sub-pixel
noise, gross outliers) and, for each problem, the inlier set and the pose returned by exactly the
calls the reference makes: cv2.solvePnPRansac(flags=SOLVEPNP_ITERATIVE) + cv2.solvePnPRefineLM.
"""
import os

import cv2
import numpy as np

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_pnp_v1.npz")
ITERS, THRESH, CONF = 400, 10.0, 0.99          # configs/infer/lmo.json:19-20, scripts/infer.py:88
CASES = [(300, 0.5, 0.3), (300, 0.3, 0.5), (120, 0.6, 0.2), (40, 0.25, 0.3), (12, 0.0, 0.1), (300, 0.7, 0.3),
         (257, 0.4, 1.0), (64, 0.5, 0.0)]       # (correspondences, outlier fraction, pixel noise sigma)
M = 300


def main() -> None:
    rng = np.random.default_rng(2024)
    K4 = np.array([600.0, 610.0, 205.0, 215.0])
    K = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1]])
    P = len(CASES)
    c2d = np.zeros((P, M, 2), np.float32)
    c3d = np.zeros((P, M, 3), np.float32)
    counts = np.zeros(P, np.int32)
    cv_success = np.zeros(P, np.int32)
    cv_R = np.zeros((P, 3, 3))
    cv_t = np.zeros((P, 3))
    cv_mask = np.zeros((P, M), np.uint8)
    gt_R = np.zeros((P, 3, 3))
    gt_t = np.zeros((P, 3))
    for p, (n, out_frac, noise) in enumerate(CASES):
        X = rng.normal(size=(n, 3)) * 60.0
        w = rng.normal(size=3)
        w = w / np.linalg.norm(w) * rng.uniform(0, np.pi)
        R = cv2.Rodrigues(w)[0]
        t = np.array([rng.uniform(-50, 50), rng.uniform(-50, 50), rng.uniform(500, 900)])
        Xc = X @ R.T + t
        x = np.stack([K4[0] * Xc[:, 0] / Xc[:, 2] + K4[2], K4[1] * Xc[:, 1] / Xc[:, 2] + K4[3]], 1)
        x += rng.normal(size=(n, 2)) * noise
        n_out = int(n * out_frac)
        oi = rng.permutation(n)[:n_out]
        # gross outliers: far (> 40 px) from the true projection, so no hypothesis can turn them into inliers
        x[oi] += rng.choice([-1.0, 1.0], size=(n_out, 2)) * rng.uniform(40, 200, size=(n_out, 2))
        c2d[p, :n], c3d[p, :n], counts[p] = x, X, n
        X32, x32 = c3d[p, :n], c2d[p, :n]
        ok, rvec, tvec, inl = cv2.solvePnPRansac(X32, x32, K, None, iterationsCount=ITERS, reprojectionError=THRESH,
                                                 confidence=CONF, flags=cv2.SOLVEPNP_ITERATIVE)
        assert ok
        rvec, tvec = cv2.solvePnPRefineLM(X32[inl[:, 0]], x32[inl[:, 0]], K, None, rvec, tvec)
        cv_success[p] = 1
        cv_R[p], cv_t[p] = cv2.Rodrigues(rvec)[0], tvec[:, 0]
        cv_mask[p, inl[:, 0]] = 1
        gt_R[p], gt_t[p] = R, t
    np.savez_compressed(OUT, coord_2d=c2d, coord_3d=c3d, counts=counts, intrinsics=np.tile(K4, (P, 1)),
                        iters=ITERS, thresh=THRESH, conf=CONF, cv_success=cv_success, cv_R=cv_R, cv_t=cv_t,
                        cv_mask=cv_mask, gt_R=gt_R, gt_t=gt_t, cv_version=cv2.__version__)
    print("wrote", OUT, os.path.getsize(OUT), "bytes; inliers per problem:", cv_mask.sum(1).tolist())


if __name__ == "__main__":
    main()
