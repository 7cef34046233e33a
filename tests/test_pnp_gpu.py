"""fp_pnp_ransac (csrc/pnp_ransac.cu) against the CPU oracle, OpenCV's golden outputs and ground truth."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_pnp_v1.npz")
SEED = 1234


def _dev():
    return torch.device("cuda", 0)


def _run(c2d, c3d, counts, K4, iters, thresh, conf, seed=SEED, offset=0):
    from foundpose_b200.utils import pnp_util
    dev = _dev()
    res = pnp_util.estimate_poses_batched(torch.from_numpy(c2d).to(dev), torch.from_numpy(c3d).to(dev),
                                          torch.from_numpy(counts).to(dev), torch.from_numpy(K4).to(dev), iters,
                                          thresh, conf, seed=seed, problem_offset=offset)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in res.items()}


def test_golden_vs_oracle_and_opencv():
    from oracle import pnp as opnp
    g = np.load(GOLDEN)
    iters, thresh, conf = int(g["iters"]), float(g["thresh"]), float(g["conf"])
    res = _run(g["coord_2d"], g["coord_3d"], g["counts"], g["intrinsics"], iters, thresh, conf)
    for p in range(g["counts"].shape[0]):
        n = int(g["counts"][p])
        ref = opnp.pnp_ransac(g["coord_2d"][p, :n], g["coord_3d"][p, :n], g["intrinsics"][p], iters, thresh, conf,
                              True, SEED, p)
        assert bool(res["success"][p]) == ref["success"]
        # integer outputs: bit-exact against the oracle
        assert int(res["best_hyp"][p]) == ref["best_hyp"]
        assert int(res["iters_run"][p]) == ref["iters_run"]
        mask = np.zeros(g["coord_2d"].shape[1], np.uint8)
        mask[ref["inliers"]] = 1
        assert np.array_equal(res["inlier_mask"][p], mask)
        assert int(res["num_inliers"][p]) == len(ref["inliers"])
        # float64 pose: same algorithm, different summation order
        assert np.abs(res["R"][p] - ref["R"]).max() < 1e-9
        assert np.abs(res["t"][p] - ref["t"]).max() < 1e-6
        # OpenCV itself (golden): identical inlier sets, pose within 1e-3 relative
        assert np.array_equal(res["inlier_mask"][p], g["cv_mask"][p])
        assert np.abs(res["R"][p] - g["cv_R"][p]).max() < 1e-5
        assert np.abs(res["t"][p] - g["cv_t"][p]).max() < 1e-3 * np.abs(g["cv_t"][p]).max()
        # a rotation
        assert np.abs(res["R"][p] @ res["R"][p].T - np.eye(3)).max() < 1e-9


def test_edge_cases_and_ragged_counts():
    from oracle import pnp as opnp
    rng = np.random.default_rng(5)
    M = 64
    K4 = np.tile(np.array([500.0, 500.0, 210.0, 210.0]), (6, 1))
    c3d = (rng.normal(size=(6, M, 3)) * 40).astype(np.float32)
    R = opnp.rodrigues(np.array([0.3, -0.2, 0.5]))
    t = np.array([5.0, -8.0, 600.0])
    Xc = c3d.astype(np.float64) @ R.T + t
    c2d = np.stack([500.0 * Xc[..., 0] / Xc[..., 2] + 210.0, 500.0 * Xc[..., 1] / Xc[..., 2] + 210.0], -1).astype(np.float32)
    c2d[4] = rng.uniform(0, 420, size=(M, 2)).astype(np.float32)       # problem 4: pure noise
    counts = np.array([0, 3, 5, 6, M, 17], np.int32)
    res = _run(c2d, c3d, counts, K4, 200, 4.0, 0.99)
    assert res["success"].tolist()[:3] == [0, 0, 0]                     # < 4 points, and 5 < 6 inliers
    assert res["iters_run"][0] == 0 and res["iters_run"][1] == 0
    assert res["success"][3] == 1 and res["num_inliers"][3] == 6
    assert res["success"][5] == 1 and res["num_inliers"][5] == 17
    assert np.all(res["inlier_mask"][5, 17:] == 0) and np.all(res["inlier_mask"][3, 6:] == 0)
    for p in (3, 5):
        assert np.abs(res["R"][p] - R).max() < 1e-4 and np.abs(res["t"][p] - t).max() < 0.05
    for p in range(6):
        n = int(counts[p])
        ref = opnp.pnp_ransac(c2d[p, :n], c3d[p, :n], K4[p], 200, 4.0, 0.99, True, SEED, p)
        assert bool(res["success"][p]) == ref["success"]
        assert int(res["best_hyp"][p]) == ref["best_hyp"] and int(res["iters_run"][p]) == ref["iters_run"]
    # failed problems return the identity pose and an empty mask
    assert np.array_equal(res["R"][0], np.eye(3)) and not res["inlier_mask"][0].any()


def test_problem_offset_and_seed():
    g = np.load(GOLDEN)
    iters, thresh, conf = int(g["iters"]), float(g["thresh"]), float(g["conf"])
    full = _run(g["coord_2d"], g["coord_3d"], g["counts"], g["intrinsics"], iters, thresh, conf)
    part = _run(g["coord_2d"][3:5], g["coord_3d"][3:5], g["counts"][3:5], g["intrinsics"][3:5], iters, thresh, conf,
                offset=3)
    for k in ("best_hyp", "iters_run", "inlier_mask", "R", "t"):
        assert np.array_equal(part[k], full[k][3:5]), k              # sharding over ranks changes nothing
    other = _run(g["coord_2d"], g["coord_3d"], g["counts"], g["intrinsics"], iters, thresh, conf, seed=99)
    assert not np.array_equal(other["best_hyp"], full["best_hyp"])
    assert np.array_equal(other["inlier_mask"], full["inlier_mask"])  # different samples, same consensus set


def test_reference_api_mirror():
    from foundpose_b200.utils import pnp_util, structs
    g = np.load(GOLDEN)
    dev = _dev()
    n = int(g["counts"][0])
    K4 = g["intrinsics"][0]
    cam = structs.PinholePlaneCameraModel(420, 420, (K4[0], K4[1]), (K4[2], K4[3]))
    corresp = {"coord_2d": torch.from_numpy(g["coord_2d"][0, :n]).to(dev), "coord_3d": torch.from_numpy(g["coord_3d"][0, :n]).to(dev)}
    ok, R, t, inliers, quality = pnp_util.estimate_pose(corresp, cam, "opencv", 400, 10.0, 0.99, True, seed=SEED)
    assert ok and R.shape == (3, 3) and t.shape == (3, 1) and inliers.shape[1] == 1 and inliers.dtype == np.int32
    assert quality == float(len(inliers)) == float(g["cv_mask"][0].sum())
    assert np.abs(R - g["cv_R"][0]).max() < 1e-5
    with pytest.raises(ValueError):
        pnp_util.estimate_pose(corresp, cam, None, 400, 10.0, 0.99, True)
    with pytest.raises(ValueError):
        pnp_util.estimate_pose({k: v.cpu() for k, v in corresp.items()}, cam, "opencv", 400, 10.0, 0.99, True)
    best = pnp_util.select_best_poses(torch.tensor([1, 1, 0, 0, 0, 0], device=dev, dtype=torch.int32),
                                      torch.tensor([10, 10, 50, 0, 0, 0], device=dev, dtype=torch.int32), 3)
    assert best.tolist() == [0, -1]


def test_full_size_batch_properties():
    """BASELINE configs[1] shape: 64 crops x 5 templates x 300 correspondences, 400 iterations."""
    from oracle import pnp as opnp
    rng = np.random.default_rng(11)
    P, M = 320, 300
    X = (rng.normal(size=(P, M, 3)) * 60).astype(np.float32)
    K4 = np.tile(np.array([600.0, 600.0, 210.0, 210.0]), (P, 1))
    Rs = np.stack([opnp.rodrigues(rng.normal(size=3)) for _ in range(P)])
    ts = np.stack([np.array([rng.uniform(-40, 40), rng.uniform(-40, 40), rng.uniform(500, 900)]) for _ in range(P)])
    Xc = np.einsum("pij,pmj->pmi", Rs, X.astype(np.float64)) + ts[:, None, :]
    x = np.stack([600.0 * Xc[..., 0] / Xc[..., 2] + 210.0, 600.0 * Xc[..., 1] / Xc[..., 2] + 210.0], -1)
    x += rng.normal(size=x.shape) * 0.3
    out = rng.random((P, M)) < 0.5
    x[out] += rng.choice([-1.0, 1.0], size=(int(out.sum()), 2)) * rng.uniform(40, 200, size=(int(out.sum()), 2))
    res = _run(x.astype(np.float32), X, np.full(P, M, np.int32), K4, 400, 10.0, 0.99)
    assert res["success"].all()
    assert np.array_equal(res["inlier_mask"].astype(bool), ~out)          # exactly the planted inliers
    assert np.abs(res["R"] - Rs).max() < 5e-3
    assert (np.abs(res["t"] - ts).max(axis=1) < 5e-3 * np.abs(ts).max(axis=1)).all()
    assert np.array_equal(res["num_inliers"], (~out).sum(1))
