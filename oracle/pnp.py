"""CPU oracle of the coarse-pose step - TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the
product path (foundpose_b200/utils/pnp_util.py -> fp_pnp_ransac) never does.

What it restates.  The reference estimates one pose per (crop, template) with
`cv2.solvePnPRansac(..., flags=SOLVEPNP_ITERATIVE)` followed by `cv2.solvePnPRefineLM` on the RANSAC
inliers (reference utils/pnp_util.py:42-72, called from scripts/infer.py:551-577 with
pnp_ransac_iter=400, pnp_inlier_thresh=10 px, configs/infer/lmo.json:18-20).  The arithmetic lives in
the third-party dependency OpenCV (opencv-python 4.9 in the reference's conda env, 4.13 in this
image), not in /root/reference.  Its published algorithm (modules/calib3d/src/solvepnp.cpp,
ptsetreg.cpp) is:

  1. RANSAC over minimal samples, inlier <=> squared reprojection error <= thresh^2, a hypothesis
     replaces the best one when it has strictly more inliers, and the iteration budget shrinks with
     RANSACUpdateNumIters(confidence, outlier ratio, model points, budget) after every improvement;
  2. a non-linear least-squares pose on the inliers of the best hypothesis (SOLVEPNP_ITERATIVE);
  3. (FoundPose) Levenberg-Marquardt refinement on the same inliers.

OpenCV draws its samples from cv::RNG and solves 5-point samples with EPnP, so its hypotheses cannot
be reproduced bit for bit outside OpenCV.  This restatement (and the CUDA kernel, which follows it
line by line) keeps steps 1-3 but makes the random part explicit and portable: hypothesis h of
problem p samples 4 distinct correspondences with a counter-based splitmix64 stream, solves P3P
(Grunert's quartic) on the first three and keeps the solution that best reprojects the fourth -
exactly what OpenCV's own SOLVEPNP_P3P kernel does with 4 points - and steps 2+3 collapse into one
Levenberg-Marquardt minimisation of the reprojection error over the inliers started from the
winning hypothesis (both OpenCV calls minimise the same cost over the same points).

Pinning: tests/golden/make_golden_pnp.py runs cv2.solvePnPRansac + cv2.solvePnPRefineLM in the
build container on seeded synthetic correspondences (known pose, sub-pixel noise, gross outliers)
and commits cv2's inlier sets and poses; tests/test_oracle_pnp_golden.py checks that this oracle
finds the same inlier sets and the same refined pose (1e-3 relative).  All arithmetic is float64.
"""

from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np

MASK64 = (1 << 64) - 1
MODEL_POINTS = 4
MIN_INLIERS_FOR_POSE = 6   # cv2's SOLVEPNP_ITERATIVE needs 6 non-planar points; fewer -> exception -> failure
LM_MAX_ITERS = 20


def splitmix64(x: int) -> int:
    z = (x + 0x9E3779B97F4A7C15) & MASK64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
    return z ^ (z >> 31)


def sample_indices(seed: int, problem: int, hyp: int, n: int) -> List[int]:
    """4 distinct indices in [0, n): draws `ctr` = 0, 1, ... of the stream (seed, problem, hyp), duplicates redrawn."""
    base = (seed ^ ((problem * 0xD1B54A32D192ED03) & MASK64) ^ ((hyp * 0x8CB92BA72F3D8DD7) & MASK64)) & MASK64
    out: List[int] = []
    ctr = 0
    while len(out) < MODEL_POINTS:
        idx = splitmix64((base + ctr * 0x2545F4914F6CDD1D) & MASK64) % n
        ctr += 1
        if idx not in out:
            out.append(idx)
    return out


def _cbrt(x: float) -> float:
    return math.copysign(abs(x) ** (1.0 / 3.0), x)


def solve_quartic_real(c4: float, c3: float, c2: float, c1: float, c0: float) -> List[float]:
    """Real roots of c4 x^4 + ... + c0 (Ferrari through the resolvent cubic, two Newton polish steps)."""
    if c4 == 0.0 or not math.isfinite(c4):
        return []
    b, c, d, e = c3 / c4, c2 / c4, c1 / c4, c0 / c4
    # depressed quartic y^4 + p y^2 + q y + r, x = y - b/4
    p = c - 3.0 * b * b / 8.0
    q = d - b * c / 2.0 + b * b * b / 8.0
    r = e - b * d / 4.0 + b * b * c / 16.0 - 3.0 * b * b * b * b / 256.0
    # resolvent cubic z^3 + 2p z^2 + (p^2 - 4r) z - q^2 = 0; take its largest real root (>= 0)
    a2, a1, a0 = 2.0 * p, p * p - 4.0 * r, -q * q
    Q = (3.0 * a1 - a2 * a2) / 9.0
    R = (9.0 * a2 * a1 - 27.0 * a0 - 2.0 * a2 * a2 * a2) / 54.0
    D = Q * Q * Q + R * R
    if D >= 0.0:
        sd = math.sqrt(D)
        z = _cbrt(R + sd) + _cbrt(R - sd) - a2 / 3.0
    else:
        th = math.acos(max(-1.0, min(1.0, R / math.sqrt(-Q * Q * Q))))
        z = 2.0 * math.sqrt(-Q) * math.cos(th / 3.0) - a2 / 3.0
    roots: List[float] = []
    if z > 1e-300:
        s = math.sqrt(z)
        for sign in (1.0, -1.0):
            # y^2 + sign*s*y + (p + z - sign*q/s)/2 = 0
            bb = sign * s
            cc = 0.5 * (p + z - sign * q / s)
            disc = bb * bb - 4.0 * cc
            if disc >= 0.0:
                sq = math.sqrt(disc)
                roots.append(0.5 * (-bb + sq) - b / 4.0)
                roots.append(0.5 * (-bb - sq) - b / 4.0)
    else:
        # biquadratic: y^4 + p y^2 + r = 0
        disc = p * p - 4.0 * r
        if disc >= 0.0:
            sq = math.sqrt(disc)
            for y2 in (0.5 * (-p + sq), 0.5 * (-p - sq)):
                if y2 >= 0.0:
                    y = math.sqrt(y2)
                    roots.append(y - b / 4.0)
                    roots.append(-y - b / 4.0)
    out: List[float] = []
    for x in roots:
        for _ in range(2):
            f = (((c4 * x + c3) * x + c2) * x + c1) * x + c0
            df = ((4.0 * c4 * x + 3.0 * c3) * x + 2.0 * c2) * x + c1
            if df != 0.0:
                x = x - f / df
        out.append(x)
    return out


def _frame(p1: np.ndarray, p2: np.ndarray, p3: np.ndarray) -> Optional[np.ndarray]:
    e1 = p2 - p1
    n1 = math.sqrt(float(e1 @ e1))
    if n1 == 0.0:
        return None
    e1 = e1 / n1
    e3 = np.cross(e1, p3 - p1)
    n3 = math.sqrt(float(e3 @ e3))
    if n3 == 0.0:
        return None
    e3 = e3 / n3
    e2 = np.cross(e3, e1)
    return np.stack([e1, e2, e3], axis=1)


def p3p_grunert(X: np.ndarray, f: np.ndarray) -> List[Tuple[np.ndarray, np.ndarray]]:
    """All (R, t) with f_i ~ R X_i + t for three world points X (3x3 rows) and unit bearings f (3x3 rows)."""
    a2 = float((X[1] - X[2]) @ (X[1] - X[2]))
    b2 = float((X[0] - X[2]) @ (X[0] - X[2]))
    c2 = float((X[0] - X[1]) @ (X[0] - X[1]))
    if a2 == 0.0 or b2 == 0.0 or c2 == 0.0:
        return []
    ca, cb, cg = float(f[1] @ f[2]), float(f[0] @ f[2]), float(f[0] @ f[1])
    k1 = (a2 - c2) / b2
    k2 = (a2 + c2) / b2
    A4 = (k1 - 1.0) ** 2 - 4.0 * c2 / b2 * ca * ca
    A3 = 4.0 * (k1 * (1.0 - k1) * cb - (1.0 - k2) * ca * cg + 2.0 * c2 / b2 * ca * ca * cb)
    A2 = 2.0 * (k1 * k1 - 1.0 + 2.0 * k1 * k1 * cb * cb + 2.0 * (b2 - c2) / b2 * ca * ca
                - 4.0 * k2 * ca * cb * cg + 2.0 * (b2 - a2) / b2 * cg * cg)
    A1 = 4.0 * (-k1 * (1.0 + k1) * cb + 2.0 * a2 / b2 * cg * cg * cb - (1.0 - k2) * ca * cg)
    A0 = (1.0 + k1) ** 2 - 4.0 * a2 / b2 * cg * cg
    Fw = _frame(X[0], X[1], X[2])
    if Fw is None:
        return []
    sols: List[Tuple[np.ndarray, np.ndarray]] = []
    for v in solve_quartic_real(A4, A3, A2, A1, A0):
        if not (v > 0.0) or not math.isfinite(v):
            continue
        den = 2.0 * (cg - v * ca)
        if den == 0.0:
            continue
        u = ((k1 - 1.0) * v * v - 2.0 * k1 * cb * v + 1.0 + k1) / den
        if not (u > 0.0):
            continue
        d1 = 1.0 + v * v - 2.0 * v * cb
        if not (d1 > 0.0):
            continue
        s1 = math.sqrt(b2 / d1)
        P = np.stack([s1 * f[0], u * s1 * f[1], v * s1 * f[2]])
        Fc = _frame(P[0], P[1], P[2])
        if Fc is None:
            continue
        R = Fc @ Fw.T
        t = P[0] - R @ X[0]
        sols.append((R, t))
    return sols


def reproj_err2(R: np.ndarray, t: np.ndarray, K4: np.ndarray, X: np.ndarray, x: np.ndarray) -> np.ndarray:
    Xc = X @ R.T + t
    with np.errstate(divide="ignore", invalid="ignore"):
        u = K4[0] * Xc[:, 0] / Xc[:, 2] + K4[2]
        v = K4[1] * Xc[:, 1] / Xc[:, 2] + K4[3]
    e = (u - x[:, 0]) ** 2 + (v - x[:, 1]) ** 2
    return np.where(np.isfinite(e), e, np.inf)


def hypothesis_pose(seed: int, problem: int, hyp: int, X: np.ndarray, x: np.ndarray,
                    K4: np.ndarray) -> Optional[Tuple[np.ndarray, np.ndarray]]:
    n = X.shape[0]
    ids = sample_indices(seed, problem, hyp, n)
    rays = np.stack([(x[ids[:3], 0] - K4[2]) / K4[0], (x[ids[:3], 1] - K4[3]) / K4[1], np.ones(3)], axis=1)
    rays = rays / np.sqrt((rays * rays).sum(axis=1, keepdims=True))
    best = None
    best_e = math.inf
    for R, t in p3p_grunert(X[ids[:3]], rays):
        e = float(reproj_err2(R, t, K4, X[ids[3]:ids[3] + 1], x[ids[3]:ids[3] + 1])[0])
        if e < best_e:
            best_e, best = e, (R, t)
    return best


def ransac_update_num_iters(p: float, ep: float, model_points: int, max_iters: int) -> int:
    """OpenCV RANSACUpdateNumIters (modules/calib3d/src/ptsetreg.cpp)."""
    p = min(max(p, 0.0), 1.0)
    ep = min(max(ep, 0.0), 1.0)
    num = max(1.0 - p, 2.2250738585072014e-308)
    denom = 1.0 - (1.0 - ep) ** model_points
    if denom < 2.2250738585072014e-308:
        return 0
    num = math.log(num)
    denom = math.log(denom)
    if denom >= 0 or -num >= max_iters * (-denom):
        return max_iters
    return int(round(num / denom))   # cvRound: nearest, ties to even - same as Python's round()


def rodrigues(w: np.ndarray) -> np.ndarray:
    th = math.sqrt(float(w @ w))
    Kx = np.array([[0.0, -w[2], w[1]], [w[2], 0.0, -w[0]], [-w[1], w[0], 0.0]])
    if th < 1e-12:
        return np.eye(3) + Kx
    return np.eye(3) + math.sin(th) / th * Kx + (1.0 - math.cos(th)) / (th * th) * (Kx @ Kx)


def _normal_equations(R, t, K4, X, x):
    Xc = X @ R.T + t
    iz = 1.0 / Xc[:, 2]
    u = K4[0] * Xc[:, 0] * iz + K4[2]
    v = K4[1] * Xc[:, 1] * iz + K4[3]
    r = np.stack([u - x[:, 0], v - x[:, 1]], axis=1)
    n = X.shape[0]
    J = np.zeros((n, 2, 6))
    # d(u,v)/dXc
    du = np.stack([K4[0] * iz, np.zeros(n), -K4[0] * Xc[:, 0] * iz * iz], axis=1)
    dv = np.stack([np.zeros(n), K4[1] * iz, -K4[1] * Xc[:, 1] * iz * iz], axis=1)
    # Xc' = Xc + dw x Xc + dt  ->  dXc/ddw = -[Xc]x
    for row, d in ((0, du), (1, dv)):
        J[:, row, 0] = d[:, 2] * Xc[:, 1] - d[:, 1] * Xc[:, 2]
        J[:, row, 1] = d[:, 0] * Xc[:, 2] - d[:, 2] * Xc[:, 0]
        J[:, row, 2] = d[:, 1] * Xc[:, 0] - d[:, 0] * Xc[:, 1]
        J[:, row, 3:6] = d
    Jf = J.reshape(-1, 6)
    rf = r.reshape(-1)
    return Jf.T @ Jf, Jf.T @ rf, float(rf @ rf)


def refine_lm(R: np.ndarray, t: np.ndarray, K4: np.ndarray, X: np.ndarray, x: np.ndarray,
              max_iters: int = LM_MAX_ITERS) -> Tuple[np.ndarray, np.ndarray]:
    """Levenberg-Marquardt on the reprojection error (left-multiplied rotation increments)."""
    lam = 1e-3
    H, g, cost = _normal_equations(R, t, K4, X, x)
    for _ in range(max_iters):
        A = H + lam * np.diag(np.diag(H))
        try:
            delta = -np.linalg.solve(A, g)
        except np.linalg.LinAlgError:
            lam *= 10.0
            continue
        Rn = rodrigues(delta[:3]) @ R
        tn = rodrigues(delta[:3]) @ t + delta[3:]
        Hn, gn, cost_n = _normal_equations(Rn, tn, K4, X, x)
        if math.isfinite(cost_n) and cost_n < cost:
            small = cost - cost_n <= 1e-12 * cost
            R, t, H, g, cost = Rn, tn, Hn, gn, cost_n
            lam = max(lam * 0.1, 1e-12)
            if small:
                break
        else:
            lam *= 10.0
            if lam > 1e12:
                break
    return R, t


def pnp_ransac(coord_2d: np.ndarray, coord_3d: np.ndarray, K4: np.ndarray, iters: int, thresh: float,
               confidence: float, refine: bool, seed: int, problem: int) -> Dict[str, object]:
    """One problem.  Returns success, R (3x3), t (3,), inliers (sorted ids), num_iters_run, best_hyp."""
    x = np.asarray(coord_2d, dtype=np.float64)
    X = np.asarray(coord_3d, dtype=np.float64)
    K4 = np.asarray(K4, dtype=np.float64)
    n = X.shape[0]
    fail = {"success": False, "R": np.eye(3), "t": np.zeros(3), "inliers": np.zeros(0, np.int64),
            "iters_run": 0, "best_hyp": -1}
    if n < MODEL_POINTS:
        return fail
    t2 = float(thresh) * float(thresh)
    best_count, best_h, best_pose = 0, -1, None
    niters = iters
    h = 0
    while h < niters:
        pose = hypothesis_pose(seed, problem, h, X, x, K4)
        if pose is not None:
            count = int((reproj_err2(pose[0], pose[1], K4, X, x) <= t2).sum())
            if count > max(best_count, MODEL_POINTS - 1):
                best_count, best_h, best_pose = count, h, pose
                niters = ransac_update_num_iters(confidence, (n - count) / n, MODEL_POINTS, niters)
        h += 1
    if best_pose is None or best_count < MIN_INLIERS_FOR_POSE:
        fail["iters_run"] = h
        return fail
    R, t = best_pose
    inl = np.nonzero(reproj_err2(R, t, K4, X, x) <= t2)[0]
    # `refine` (FoundPose's pnp_refine_lm) is accepted for signature parity: solvePnPRansac already ends with a
    # non-linear solve on the inliers and solvePnPRefineLM minimises the same cost over the same points.
    R, t = refine_lm(R, t, K4, X[inl], x[inl])
    return {"success": True, "R": R, "t": t, "inliers": inl.astype(np.int64), "iters_run": h, "best_hyp": best_h}
