"""Coarse pose from 2D-3D correspondences on the GPU - drop-in for the reference's utils/pnp_util.py.

`estimate_pose` keeps the reference signature and return tuple (utils/pnp_util.py:20-84): the
reference runs `cv2.solvePnPRansac(flags=SOLVEPNP_ITERATIVE)` + `cv2.solvePnPRefineLM` on the host
for one (crop, template) at a time (scripts/infer.py:551-577); here the same step is one launch of
`fp_pnp_ransac` (csrc/pnp_ransac.cu) and `estimate_poses_batched` runs every (crop, template) pair
of a batch at once, straight from the device buffers `pipeline.RetrievalEngine.match` leaves behind.

What is identical to OpenCV and what is restated is written down in oracle/pnp.py: the RANSAC
semantics (inlier test, strict improvement, RANSACUpdateNumIters), the final non-linear least
squares on the inliers and the >= 6 inliers rule are kept; cv::RNG sampling and the EPnP minimal
solver are replaced by an explicit splitmix64 stream and P3P + a fourth point, so a fixed `seed`
makes the result reproducible across runs, devices and the CPU oracle.
"""

from __future__ import annotations

from typing import Any, Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from foundpose_b200 import _native
from foundpose_b200.utils import logging, misc, structs

logger: logging.Logger = logging.get_logger()

DEFAULT_SEED = 0


def get_intrinsics_vector(camera: structs.CameraModel) -> np.ndarray:
    """(fx, fy, cx, cy) of the camera - the entries of misc.get_intrinsic_matrix (reference misc.py:325-341)."""
    return np.array([camera.f[0], camera.f[1], camera.c[0], camera.c[1]], dtype=np.float64)


def estimate_poses_batched(coord_2d: torch.Tensor, coord_3d: torch.Tensor, counts: torch.Tensor,
                           intrinsics: torch.Tensor, pnp_ransac_iter: int, pnp_inlier_thresh: float,
                           pnp_required_ransac_conf: float, seed: int = DEFAULT_SEED,
                           problem_offset: int = 0) -> Dict[str, torch.Tensor]:
    """P problems at once, everything on the device.

    coord_2d fp32 [P,M,2], coord_3d fp32 [P,M,3], counts int32 [P], intrinsics fp64 [P,4] (fx,fy,cx,cy).
    Returns device tensors: success int32 [P], R fp64 [P,3,3], t fp64 [P,3] (model -> camera),
    inlier_mask uint8 [P,M], num_inliers / iters_run / best_hyp int32 [P].  Problems with fewer than
    6 correspondences fail like the reference skips them (scripts/infer.py:558-561).
    """
    if not coord_2d.is_cuda:
        raise ValueError("foundpose_b200 runs on CUDA tensors only (no CPU fallback)")
    return _native.pnp_ransac(coord_2d.contiguous(), coord_3d.contiguous(), counts.contiguous(),
                              intrinsics.contiguous(), int(pnp_ransac_iter), float(pnp_inlier_thresh),
                              float(pnp_required_ransac_conf), int(seed), int(problem_offset))


def estimate_pose(
    corresp: Dict[str, Any],
    camera_c2w: structs.PinholePlaneCameraModel,
    pnp_type: str,
    pnp_ransac_iter: int,
    pnp_inlier_thresh: float,
    pnp_required_ransac_conf: float,
    pnp_refine_lm: bool,
    seed: int = DEFAULT_SEED,
) -> Tuple[bool, Optional[np.ndarray], Optional[np.ndarray], Optional[np.ndarray], Optional[float]]:
    """Same contract as the reference's estimate_pose: (success, R_m2c 3x3, t_m2c 3x1, inliers Nx1, quality).

    `pnp_type` must be "opencv" (the only value the reference accepts, utils/pnp_util.py:40,81-82: it
    names the algorithm family, executed here by the CUDA kernel).  `pnp_refine_lm` is accepted for
    signature parity: the kernel always ends with the Levenberg-Marquardt solve on the inliers that
    both solvePnPRansac and solvePnPRefineLM perform.
    """
    if pnp_type != "opencv":
        raise ValueError("Unsupported PnP type")
    c2d = corresp["coord_2d"]
    c3d = corresp["coord_3d"]
    if not isinstance(c2d, torch.Tensor):
        c2d, c3d = misc.array_to_tensor(np.asarray(c2d)), misc.array_to_tensor(np.asarray(c3d))
    if not c2d.is_cuda:
        raise ValueError("foundpose_b200 runs on CUDA tensors only (no CPU fallback)")
    dev = c2d.device
    n = int(c2d.shape[0])
    if n == 0:
        return False, None, None, None, None
    res = estimate_poses_batched(
        c2d.to(torch.float32).reshape(1, n, 2), c3d.to(torch.float32).reshape(1, n, 3),
        torch.full((1,), n, dtype=torch.int32, device=dev),
        torch.from_numpy(get_intrinsics_vector(camera_c2w)).to(dev).reshape(1, 4),
        pnp_ransac_iter, pnp_inlier_thresh, pnp_required_ransac_conf, seed)
    ok = bool(res["success"][0].item())
    if not ok:
        # cv2 returns success=False with whatever rvec/tvec it had; the reference then stores nothing.
        return False, np.eye(3), np.zeros((3, 1)), None, 0.0
    inliers = torch.nonzero(res["inlier_mask"][0]).to(torch.int32).cpu().numpy()      # N x 1, like cv2
    return True, res["R"][0].cpu().numpy(), res["t"][0].cpu().numpy().reshape(3, 1), inliers, float(len(inliers))


def select_best_poses(success: torch.Tensor, num_inliers: torch.Tensor, top_n: int) -> torch.Tensor:
    """Index of the best template per crop: max quality = inlier count, first wins ties
    (scripts/infer.py:592-603).  success / num_inliers are [B*top_n]; returns int64 [B], -1 = no pose."""
    q = torch.where(success.bool(), num_inliers, torch.full_like(num_inliers, -1)).reshape(-1, top_n)
    best = torch.argmax(q, dim=1)            # argmax returns the first maximal index
    return torch.where(q.gather(1, best[:, None])[:, 0] >= 0, best, torch.full_like(best, -1))
