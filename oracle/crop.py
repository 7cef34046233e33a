"""CPU restatement of the crop stage that feeds the hot path.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs; the
product path never touches it.

Follows scripts/infer.py:411-462 (crop box -> virtual camera -> warp image + mask -> box of the warped
mask) with utils/misc.py:171-205 (calc_crop_box), :208-277 (construct_crop_camera), :458-519
(warp_image), utils/geometry.py:52-88 (gen_look_at_matrix), :129-146 (from_two_vectors) and
utils/structs.py:477-500 (pinhole world/eye/window maps).  `cv2.remap` is a third-party dependency
(OpenCV 4.x, imgproc/src/imgwarp.cpp `remap`); its published algorithm is restated in
`remap_linear_f32` / `remap_nearest`:

  * float maps are converted to fixed point with 5 fractional bits, `cvRound(x * 32)` (round half to
    even), integer part `>> 5` saturated to int16, fraction `& 31`;
  * INTER_LINEAR on float32 pixels: the four taps are weighted with the float table
    (1-fy)(1-fx), (1-fy)fx, fy(1-fx), fy*fx (all exact in fp32) and summed left to right in fp32;
    taps outside the image take the BORDER_CONSTANT value 0;
  * INTER_AREA is treated as INTER_LINEAR by remap;
  * INTER_NEAREST rounds the float map with `cvRound` (saturated to int16) and copies the pixel, 0 outside.

Pinned by tests/golden/golden_crop_v1.pt: outputs of the reference's own utils/misc.py +
utils/structs.py + utils/geometry.py and of cv2.remap (OpenCV 4.13) run in the build container
(tests/golden/make_golden_crop.py).
"""

from __future__ import annotations

from typing import Dict, Tuple

import numpy as np


def calc_crop_box(box_ltrb, box_scaling_factor: float = 1.0, make_square: bool = False) -> Tuple[float, float, float, float]:
    """utils/misc.py:171-205.  `box_ltrb` = (left, top, right, bottom)."""
    left, top, right, bottom = [float(v) for v in box_ltrb]
    width, height = right - left, bottom - top
    cw, ch = width * box_scaling_factor, height * box_scaling_factor
    if make_square:
        cw = ch = max(cw, ch)
    x_pad, y_pad = 0.5 * (cw - width), 0.5 * (ch - height)
    return (left - x_pad, top - y_pad, right + x_pad, bottom + y_pad)


def _normalized(v: np.ndarray, axis: int = -1) -> np.ndarray:
    return v / np.linalg.norm(v, axis=axis, keepdims=True)


def _from_two_vectors(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """utils/geometry.py:129-146 (Rodrigues form of the rotation taking a to b)."""
    a, b = _normalized(a), _normalized(b)
    v = np.cross(a, b)
    s, c = np.linalg.norm(v), np.dot(a, b)
    vm = np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]], dtype=v.dtype)
    return np.eye(3, dtype=a.dtype) + vm + (vm @ vm) * (1 - c) / max(s * s, 1e-15)


def construct_crop_camera(box_ltrb, f, c, T_world_from_eye: np.ndarray, viewport_size: Tuple[int, int],
                          viewport_rel_pad: float) -> Dict[str, object]:
    """utils/misc.py:208-277.  Returns {"width","height","f","c","T_world_from_eye"} of the virtual camera.

    f and c of the result are float32 pairs, as in the reference (it builds them from float32 arrays).
    """
    left, top, right, bottom = [float(v) for v in box_ltrb]
    T = np.asarray(T_world_from_eye, dtype=np.float64)
    fm = 0.5 * (f[0] + f[1])
    cx, cy = c
    corners = np.array([[left - cx, top - cy, fm], [right - cx, top - cy, fm],
                        [left - cx, bottom - cy, fm], [right - cx, bottom - cy, fm]], dtype=np.float64)
    corners /= np.linalg.norm(corners, axis=1, keepdims=True)
    centroid_c = corners.mean(axis=0)
    centroid_w = T.dot(np.hstack([centroid_c, 1]).reshape(4, 1))[:3, 0]
    radius = np.linalg.norm(corners - centroid_c, axis=1).max()

    # gen_look_at_matrix (geometry.py:52-88): rotate the camera so +z passes through the centroid.
    w2c = np.linalg.inv(T)
    center_local = w2c[:3, :3] @ centroid_w + w2c[:3, 3]
    delta = _from_two_vectors(np.array([0, 0, 1], dtype=centroid_w.dtype), center_local / np.linalg.norm(center_local))
    c2w_new = np.linalg.inv(w2c).copy()
    c2w_new[:3, :3] = c2w_new[:3, :3] @ delta
    c2w_new[:3, :3] = c2w_new[:3, :3] @ np.eye(3)          # roll angle 0
    w2vc = np.linalg.inv(c2w_new)

    centroid_vc = (w2vc.dot(np.hstack([centroid_w, 1]).reshape(4, 1))[:3, :].T).squeeze()
    f_orig = np.array(f, dtype=np.float32)
    radius_2d = f_orig * radius / centroid_vc[2]
    extent_2d = (1.0 + viewport_rel_pad) * radius_2d
    cx_cy = np.array(viewport_size, dtype=np.float32) / 2.0 - 0.5
    fx_fy = f_orig * cx_cy / extent_2d
    return {"width": int(viewport_size[0]), "height": int(viewport_size[1]), "f": tuple(fx_fy), "c": tuple(cx_cy),
            "T_world_from_eye": np.linalg.inv(w2vc)}


def warp_map(src_f, src_c, src_T: np.ndarray, dst: Dict[str, object], depth_check: bool = True
             ) -> Tuple[np.ndarray, np.ndarray]:
    """utils/misc.py:493-516: float32 (map_x, map_y) [H, W] of source window coordinates."""
    W, H = dst["width"], dst["height"]
    px, py = np.meshgrid(np.arange(W), np.arange(H))
    win = np.column_stack((px.flatten(), py.flatten()))
    q = (np.asarray(win) - dst["c"]) / dst["f"]                          # structs.py:496-500
    v = np.stack((q[:, 0], q[:, 1], np.ones_like(q[:, 0])), axis=-1)
    v = _normalized(v, axis=-1)
    Td = np.asarray(dst["T_world_from_eye"])
    world = v @ Td[:3, :3].T + Td[:3, 3]                                  # structs.py:485-489
    Ts = np.asarray(src_T)
    eye = (world - Ts[:3, 3]) @ Ts[:3, :3]                                # structs.py:477-483 (rotate by R^T)
    win_s = eye[:, :2] / eye[:, 2, None] * src_f + src_c                  # structs.py:491-494
    if depth_check:
        win_s[eye[:, 2] < 0] = -1
    win_s = win_s.astype(np.float32)
    return win_s[:, 0].reshape(H, W), win_s[:, 1].reshape(H, W)


def _fixed_point(map_xy: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    s = np.rint(map_xy.astype(np.float32) * np.float32(32)).astype(np.int64)   # cvRound = round half to even
    return np.clip(s >> 5, -32768, 32767), s & 31


def remap_linear_f32(src: np.ndarray, map_x: np.ndarray, map_y: np.ndarray) -> np.ndarray:
    """cv2.remap(src float32 [H,W] or [H,W,C], map_x, map_y, INTER_LINEAR), BORDER_CONSTANT 0."""
    assert src.dtype == np.float32
    img = src if src.ndim == 3 else src[:, :, None]
    h, w, _ = img.shape
    ix, fx = _fixed_point(map_x)
    iy, fy = _fixed_point(map_y)
    ax, ay = (fx.astype(np.float32) / np.float32(32)), (fy.astype(np.float32) / np.float32(32))
    one = np.float32(1)
    weights = [(one - ay) * (one - ax), (one - ay) * ax, ay * (one - ax), ay * ax]

    def tap(yy, xx):
        inside = (xx >= 0) & (xx < w) & (yy >= 0) & (yy < h)
        vals = img[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)]
        return np.where(inside[..., None], vals, np.float32(0))

    taps = [tap(iy, ix), tap(iy, ix + 1), tap(iy + 1, ix), tap(iy + 1, ix + 1)]
    out = taps[0] * weights[0][..., None]
    for t, wt in zip(taps[1:], weights[1:]):
        out = out + t * wt[..., None]                     # fp32, left to right, no fused multiply-add
    out = out.astype(np.float32)
    return out if src.ndim == 3 else out[:, :, 0]


def remap_nearest(src: np.ndarray, map_x: np.ndarray, map_y: np.ndarray) -> np.ndarray:
    """cv2.remap(src, map_x, map_y, INTER_NEAREST), BORDER_CONSTANT 0."""
    h, w = src.shape[:2]
    ix = np.clip(np.rint(map_x.astype(np.float32)).astype(np.int64), -32768, 32767)
    iy = np.clip(np.rint(map_y.astype(np.float32)).astype(np.int64), -32768, 32767)
    inside = (ix >= 0) & (ix < w) & (iy >= 0) & (iy < h)
    vals = src[np.clip(iy, 0, h - 1), np.clip(ix, 0, w - 1)]
    if src.ndim == 3:
        inside = inside[..., None]
    return np.where(inside, vals, np.zeros((), dtype=src.dtype)).astype(src.dtype)


def calc_2d_box_of_mask(mask: np.ndarray) -> np.ndarray:
    """infer.py:449-456 + utils/misc.py:279-310: (x1, y1, x2, y2) of the non-zero pixels, zeros when empty."""
    ys, xs = mask.nonzero()
    if len(xs) == 0:
        return np.zeros(4, dtype=np.float32)
    return np.array([xs.min(), ys.min(), xs.max(), ys.max()], dtype=np.float32)


def crop_instance(image_u8_hwc: np.ndarray, mask_u8: np.ndarray, box_ltrb, f, c, T_world_from_eye: np.ndarray,
                  crop_size: Tuple[int, int], crop_rel_pad: float) -> Dict[str, object]:
    """scripts/infer.py:396-459 for one instance: float image in [0,1] HWC, warped mask, new box, crop camera."""
    image = image_u8_hwc.astype(np.float32) / 255.0
    crop_box = calc_crop_box(box_ltrb, make_square=True)
    cam = construct_crop_camera(crop_box, f, c, T_world_from_eye, crop_size, crop_rel_pad)
    map_x, map_y = warp_map(np.asarray(f), np.asarray(c), T_world_from_eye, cam)
    # INTER_AREA when shrinking, INTER_LINEAR otherwise: the same code path inside remap.
    warped = remap_linear_f32(image, map_x, map_y)
    warped_mask = remap_nearest(mask_u8, map_x, map_y)
    return {"image": warped, "mask": warped_mask, "box": calc_2d_box_of_mask(warped_mask), "camera": cam,
            "map_x": map_x, "map_y": map_y}
