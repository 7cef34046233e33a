"""The helpers of the reference's utils/misc.py on or next to the hot path: Timer and array<->tensor
(:30-45, 559-603) and the crop stage (:171-277, 458-519)."""

import time
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from foundpose_b200 import _native
from foundpose_b200.utils import logging, structs

logger: logging.Logger = logging.get_logger()


class Timer:
    """Wall-clock timer (reference utils/misc.py:30-45).

    Unlike the reference, `elapsed` synchronises the current CUDA device first when
    `cuda_sync=True`, so stage times are attributed to the stage that launched the work.
    """

    def __init__(self, enabled: bool = True, cuda_sync: bool = False) -> None:
        self.enabled = enabled
        self.cuda_sync = cuda_sync
        self.start_time = None

    def start(self):
        if self.enabled:
            if self.cuda_sync and torch.cuda.is_available():
                torch.cuda.synchronize()
            self.start_time = time.time()

    def elapsed(self, msg="Elapsed") -> Optional[float]:
        if self.enabled:
            if self.cuda_sync and torch.cuda.is_available():
                torch.cuda.synchronize()
            elapsed = time.time() - self.start_time
            logger.info(f"{msg}: {elapsed:.5f}s")
            return elapsed
        else:
            return None


def array_to_tensor(array: np.ndarray, make_array_writeable: bool = True) -> torch.Tensor:
    if not array.flags.writeable:
        if make_array_writeable and array.flags.owndata:
            array.setflags(write=True)
        else:
            array = np.array(array)
    return torch.from_numpy(array)


def tensor_to_array(tensor: torch.Tensor) -> np.ndarray:
    return tensor.detach().cpu().numpy()


# ---------------------------------------------------------------------------------------------
# Crop stage (reference utils/misc.py:171-277, 458-519; scripts/infer.py:411-459).  The box and camera
# geometry is a handful of scalars per instance and stays on the host in fp64 like the reference; all
# per-pixel work (both warps, the [0,1] scaling, the HWC->CHW transpose and the box of the warped mask)
# is one launch of csrc/crop_warp.cu for the whole batch of instances.
# ---------------------------------------------------------------------------------------------
INTER_NEAREST, INTER_LINEAR, INTER_AREA = 0, 1, 3          # OpenCV's codes, so callers can pass cv2.INTER_*


def calc_crop_box(box: structs.AlignedBox2f, box_scaling_factor: float = 1.0,
                  make_square: bool = False) -> structs.AlignedBox2f:
    """Scales a box about its centre and optionally makes it square (reference :171-205)."""
    side_w, side_h = box.width * box_scaling_factor, box.height * box_scaling_factor
    if make_square:
        side_w = side_h = max(side_w, side_h)
    grow_x, grow_y = 0.5 * (side_w - box.width), 0.5 * (side_h - box.height)
    return structs.AlignedBox2f(box.left - grow_x, box.top - grow_y, box.right + grow_x, box.bottom + grow_y)


def construct_crop_cameras(boxes_ltrb: np.ndarray, camera_model_c2w: structs.CameraModel,
                           viewport_size: Tuple[int, int], viewport_rel_pad: float
                           ) -> List[structs.PinholePlaneCameraModel]:
    """construct_crop_camera (reference :208-277) for all instances of one image at once.

    The virtual camera keeps the optical centre, looks at the centroid of the four unit rays through
    the box corners, and gets the focal length that makes the corner sphere (+ padding) fill the viewport.
    """
    boxes = np.asarray(boxes_ltrb, dtype=np.float64).reshape(-1, 4)
    T = camera_model_c2w.T_world_from_eye
    R, t = T[:3, :3], T[:3, 3]
    f_mean = 0.5 * (camera_model_c2w.f[0] + camera_model_c2w.f[1])
    cx, cy = camera_model_c2w.c
    xs = boxes[:, [0, 2, 0, 2]] - cx                                  # corner order: lt, rt, lb, rb
    ys = boxes[:, [1, 1, 3, 3]] - cy
    rays = np.stack([xs, ys, np.full_like(xs, f_mean)], axis=-1)      # [B, 4, 3]
    rays /= np.linalg.norm(rays, axis=-1, keepdims=True)
    centroid = rays.mean(axis=1)                                      # [B, 3] in the camera frame
    radius = np.linalg.norm(rays - centroid[:, None, :], axis=-1).max(axis=1)
    # look-at: the rotation taking +z onto the centroid direction (geometry.py:52-88, 129-146)
    d = centroid / np.linalg.norm(centroid, axis=-1, keepdims=True)
    axis = np.stack([-d[:, 1], d[:, 0], np.zeros(len(d))], axis=-1)   # z x d
    sin2 = np.maximum((axis ** 2).sum(-1), 1e-15)
    K = np.zeros((len(d), 3, 3))
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0] = -axis[:, 2], axis[:, 1], axis[:, 2]
    K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -axis[:, 0], -axis[:, 1], axis[:, 0]
    delta = np.eye(3) + K + (K @ K) * ((1.0 - d[:, 2]) / sin2)[:, None, None]
    R_new = R @ delta                                                 # [B, 3, 3]
    depth = np.einsum("bij,bi->bj", delta, centroid)[:, 2]            # centroid z in the virtual camera
    f_orig = np.array(camera_model_c2w.f, dtype=np.float32)
    half = np.array(viewport_size, dtype=np.float32) / 2.0 - 0.5
    cameras = []
    for b in range(len(boxes)):
        extent = (1.0 + viewport_rel_pad) * (f_orig * radius[b] / depth[b])
        T_new = np.eye(4)
        T_new[:3, :3], T_new[:3, 3] = R_new[b], t
        cameras.append(structs.PinholePlaneCameraModel(width=viewport_size[0], height=viewport_size[1],
                                                       f=tuple(f_orig * half / extent), c=tuple(half),
                                                       T_world_from_eye=T_new))
    return cameras


def construct_crop_camera(box: structs.AlignedBox2f, camera_model_c2w: structs.CameraModel,
                          viewport_size: Tuple[int, int], viewport_rel_pad: float) -> structs.CameraModel:
    """Single-instance form with the reference's signature (:208-277)."""
    return construct_crop_cameras(box.array_ltrb()[None], camera_model_c2w, viewport_size, viewport_rel_pad)[0]


def pack_crop_params(src_cameras: Sequence[structs.CameraModel], dst_cameras: Sequence[structs.CameraModel],
                     image_index: Optional[Sequence[int]] = None) -> np.ndarray:
    """fp64 [B, 40] parameter block of fp_crop_warp (layout in include/foundpose_b200.h)."""
    n = len(dst_cameras)
    prm = np.zeros((n, _native.CROP_PARAM_STRIDE), dtype=np.float64)
    for b, (src, dst) in enumerate(zip(src_cameras, dst_cameras)):
        prm[b, 0:2], prm[b, 2:4] = dst.f, dst.c
        prm[b, 4:13], prm[b, 13:16] = dst.T_world_from_eye[:3, :3].reshape(9), dst.T_world_from_eye[:3, 3]
        prm[b, 16:25], prm[b, 25:28] = src.T_world_from_eye[:3, :3].reshape(9), src.T_world_from_eye[:3, 3]
        prm[b, 28:30], prm[b, 30:32] = src.f, src.c
        prm[b, 32] = 0 if image_index is None else image_index[b]
    return prm


def warp_crops(images: Optional[torch.Tensor], masks: Optional[torch.Tensor],
               src_cameras: Sequence[structs.CameraModel], dst_cameras: Sequence[structs.CameraModel],
               image_index: Optional[Sequence[int]] = None):
    """Warps B instances into their crop cameras in one launch.

    images: CUDA [n_img, H, W, C] uint8 (scaled to [0,1]) or fp32; masks: CUDA uint8 [B, H, W].
    Returns (crops fp32 [B, C, h, w], warped masks uint8 [B, h, w], boxes fp32 [B, 4]); see fp_crop_warp.
    """
    ref = images if images is not None else masks
    if ref is None or not ref.is_cuda:
        raise ValueError("foundpose_b200 runs on CUDA tensors only (no CPU fallback)")
    sizes = {(cam.width, cam.height) for cam in dst_cameras}
    assert len(sizes) == 1, "all crop cameras of a batch share one viewport"
    crop_w, crop_h = sizes.pop()
    prm = torch.from_numpy(pack_crop_params(src_cameras, dst_cameras, image_index)).to(ref.device, non_blocking=True)
    return _native.crop_warp(images, masks, prm, int(crop_w), int(crop_h))


def warp_image(src_camera: structs.CameraModel, dst_camera: structs.CameraModel, src_image,
               interpolation: int = INTER_LINEAR, depth_check: bool = True, factor_to_downsample: int = 1):
    """Single-image form with the reference's signature (:458-519); numpy in -> numpy out, tensor -> tensor.

    INTER_LINEAR / INTER_AREA take fp32 or uint8 [H,W] / [H,W,C] images (uint8 is NOT rescaled here,
    it is converted to fp32 like cv2 would interpolate it only for fp32 inputs - pass fp32 for exact
    cv2 parity); INTER_NEAREST takes uint8 [H,W] masks.
    """
    if not depth_check or factor_to_downsample != 1:
        raise NotImplementedError("warp_image: only the configuration used by scripts/infer.py is provided")
    as_numpy = isinstance(src_image, np.ndarray)
    img = torch.from_numpy(np.ascontiguousarray(src_image)) if as_numpy else src_image
    img = img.cuda().contiguous()
    if interpolation == INTER_NEAREST:
        if img.dtype != torch.uint8 or img.dim() != 2:
            raise NotImplementedError("warp_image(INTER_NEAREST) takes uint8 [H, W] masks")
        out = warp_crops(None, img[None], [src_camera], [dst_camera])[1][0]
    elif interpolation in (INTER_LINEAR, INTER_AREA):
        if img.dtype != torch.float32:
            raise NotImplementedError("warp_image(INTER_LINEAR) takes float32 images, as scripts/infer.py passes them")
        planar = warp_crops((img if img.dim() == 3 else img[:, :, None])[None], None, [src_camera], [dst_camera])[0][0]
        out = planar.permute(1, 2, 0).contiguous() if img.dim() == 3 else planar[0]
    else:
        raise ValueError(f"Unsupported interpolation {interpolation}")
    return out.cpu().numpy() if as_numpy else out


def crop_instances(image_u8_hwc: torch.Tensor, masks_u8: torch.Tensor, boxes_ltrb: np.ndarray,
                   camera_c2w: structs.CameraModel, crop_size: Tuple[int, int], crop_rel_pad: float):
    """The `opts.crop` branch of scripts/infer.py:416-459 for all instances of one image.

    image_u8_hwc: CUDA uint8 [H, W, 3]; masks_u8: CUDA uint8 [B, H, W] modal masks; boxes_ltrb: [B, 4]
    amodal boxes.  Returns (crops fp32 [B,3,h,w] in [0,1], warped masks uint8 [B,h,w], boxes of the warped
    masks fp32 [B,4] on the device, list of crop cameras).
    """
    boxes = np.asarray(boxes_ltrb, dtype=np.float64).reshape(-1, 4)
    crop_boxes = np.stack([calc_crop_box(structs.AlignedBox2f(*bx), make_square=True).array_ltrb() for bx in boxes])
    cameras = construct_crop_cameras(crop_boxes, camera_c2w, crop_size, crop_rel_pad)
    crops, warped_masks, new_boxes = warp_crops(image_u8_hwc[None], masks_u8, [camera_c2w] * len(cameras), cameras)
    return crops, warped_masks, new_boxes, cameras
