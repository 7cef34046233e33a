"""The coarse-pose oracle (oracle/pnp.py) against OpenCV's own outputs (tests/golden/golden_pnp_v1.npz)."""
import os

import numpy as np
import pytest

from oracle import pnp as opnp

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_pnp_v1.npz")
SEED = 1234


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def test_oracle_matches_opencv_inliers_and_pose(golden):
    P = golden["counts"].shape[0]
    for p in range(P):
        n = int(golden["counts"][p])
        res = opnp.pnp_ransac(golden["coord_2d"][p, :n], golden["coord_3d"][p, :n], golden["intrinsics"][p],
                              int(golden["iters"]), float(golden["thresh"]), float(golden["conf"]), True, SEED, p)
        assert res["success"] == bool(golden["cv_success"][p])
        mask = np.zeros(golden["cv_mask"].shape[1], np.uint8)
        mask[res["inliers"]] = 1
        # identical inlier sets: the data separate inliers (sub-pixel noise) from outliers (> 40 px)
        assert np.array_equal(mask, golden["cv_mask"][p]), f"problem {p}"
        # same least-squares optimum over the same points: 1e-3 relative is the bar, 1e-6 is what we see
        assert np.abs(res["R"] - golden["cv_R"][p]).max() < 1e-5
        assert np.abs(res["t"] - golden["cv_t"][p]).max() < 1e-3 * np.abs(golden["cv_t"][p]).max()


def test_quartic_solver_known_roots():
    # (x-1)(x-2)(x+3)(x-0.5)
    coeffs = np.poly([1.0, 2.0, -3.0, 0.5])
    roots = sorted(opnp.solve_quartic_real(*coeffs))
    assert np.allclose(roots, [-3.0, 0.5, 1.0, 2.0], atol=1e-9)
    assert opnp.solve_quartic_real(1.0, 0.0, 2.0, 0.0, 5.0) == []          # no real roots
    assert np.allclose(sorted(opnp.solve_quartic_real(1.0, 0.0, -5.0, 0.0, 4.0)), [-2, -1, 1, 2])   # biquadratic


def test_p3p_recovers_exact_pose():
    rng = np.random.default_rng(0)
    for _ in range(20):
        X = rng.normal(size=(3, 3)) * 50
        w = rng.normal(size=3)
        R = opnp.rodrigues(w)
        t = np.array([3.0, -4.0, 400.0])
        Xc = X @ R.T + t
        f = Xc / np.linalg.norm(Xc, axis=1, keepdims=True)
        sols = opnp.p3p_grunert(X, f)
        assert any(np.abs(Rs - R).max() < 1e-6 and np.abs(ts - t).max() < 1e-4 for Rs, ts in sols)


def test_sampling_is_distinct_and_reproducible():
    a = opnp.sample_indices(7, 3, 11, 5)
    assert a == opnp.sample_indices(7, 3, 11, 5) and len(set(a)) == 4 and all(0 <= i < 5 for i in a)
    assert opnp.sample_indices(7, 3, 12, 300) != opnp.sample_indices(7, 3, 11, 300)
    assert opnp.splitmix64(0) == 0xE220A8397B1DCDAF        # published splitmix64 test vector


def test_too_few_points_and_iteration_budget():
    rng = np.random.default_rng(1)
    res = opnp.pnp_ransac(rng.normal(size=(3, 2)), rng.normal(size=(3, 3)), np.array([600.0, 600, 200, 200]),
                          50, 10.0, 0.99, True, 0, 0)
    assert not res["success"] and res["iters_run"] == 0
    # OpenCV's RANSACUpdateNumIters: known values
    assert opnp.ransac_update_num_iters(0.99, 0.5, 4, 400) == 71
    assert opnp.ransac_update_num_iters(0.99, 0.0, 4, 400) == 0
    assert opnp.ransac_update_num_iters(0.99, 1.0, 4, 400) == 400


def test_refine_lm_converges_from_a_perturbed_pose():
    rng = np.random.default_rng(3)
    X = rng.normal(size=(40, 3)) * 50
    R = opnp.rodrigues(np.array([0.2, -0.4, 0.1]))
    t = np.array([10.0, -5.0, 700.0])
    K4 = np.array([600.0, 590.0, 210.0, 205.0])
    Xc = X @ R.T + t
    x = np.stack([K4[0] * Xc[:, 0] / Xc[:, 2] + K4[2], K4[1] * Xc[:, 1] / Xc[:, 2] + K4[3]], 1)
    R0 = opnp.rodrigues(np.array([0.03, -0.02, 0.04])) @ R
    t0 = t + np.array([3.0, -2.0, 15.0])
    R1, t1 = opnp.refine_lm(R0, t0, K4, X, x)
    assert np.abs(R1 - R).max() < 1e-8 and np.abs(t1 - t).max() < 1e-5
    assert np.abs(R1 @ R1.T - np.eye(3)).max() < 1e-12


def test_select_best_poses_host_logic():
    import torch

    from foundpose_b200.utils import pnp_util

    success = torch.tensor([1, 1, 1, 0, 0, 0, 1, 0, 1], dtype=torch.int32)
    inliers = torch.tensor([12, 40, 40, 99, 0, 0, 7, 50, 7], dtype=torch.int32)
    best = pnp_util.select_best_poses(success, inliers, 3)
    # first maximal quality wins (scripts/infer.py:592-603); failed problems never win; no success -> -1
    assert best.tolist() == [1, -1, 0]
