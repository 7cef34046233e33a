"""The two structures of the reference's utils/structs.py that the crop stage needs:
`AlignedBox2f` (:115-252) and the pinhole camera (`CameraModel` :255-520, `PinholePlaneCameraModel`
:672-680).  Plain host-side geometry; the per-pixel work happens in csrc/crop_warp.cu.
"""

from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np


class AlignedBox2f:
    """Axis-aligned 2D box given by its left / top / right / bottom edges (reference :115-252)."""

    def __init__(self, left: float, top: float, right: float, bottom: float):
        self._left, self._top, self._right, self._bottom = left, top, right, bottom

    def __repr__(self) -> str:
        return f"AlignedBox2f(left: {self._left}, top: {self._top}, right: {self._right}, bottom: {self._bottom})"

    left = property(lambda self: self._left)
    top = property(lambda self: self._top)
    right = property(lambda self: self._right)
    bottom = property(lambda self: self._bottom)
    width = property(lambda self: self._right - self._left)
    height = property(lambda self: self._bottom - self._top)

    def array_ltrb(self) -> np.ndarray:
        return np.array([self._left, self._top, self._right, self._bottom])

    def array_ltwh(self) -> np.ndarray:
        return np.array([self._left, self._top, self.width, self.height])

    def pad(self, width: float, height: float) -> "AlignedBox2f":
        return AlignedBox2f(self._left - width, self._top - height, self._right + width, self._bottom + height)

    def clip(self, boundary: "AlignedBox2f") -> "AlignedBox2f":
        return AlignedBox2f(max(self._left, boundary.left), max(self._top, boundary.top),
                            min(self._right, boundary.right), min(self._bottom, boundary.bottom))


class CameraModel:
    """Pinhole intrinsics (f, c in pixels) + camera-to-world extrinsics (reference :255-352, 477-500)."""

    def __init__(self, width: int, height: int, f, c: Sequence[float],
                 T_world_from_eye: Optional[np.ndarray] = None, serial: str = "") -> None:
        self.width, self.height, self.serial = width, height, serial
        self.f: Tuple[float, float] = tuple(np.broadcast_to(f, 2))
        self.c: Tuple[float, float] = tuple(c)
        if T_world_from_eye is None:
            self.T_world_from_eye = np.eye(4)
        else:
            T = np.array(T_world_from_eye, dtype=np.float64)
            if T.shape == (3, 4):
                T = np.vstack([T, [0.0, 0.0, 0.0, 1.0]])
            err = np.abs((T.T @ T)[:3, :3] - np.eye(3)).max()
            if err >= 1.0e-5:
                raise ValueError(f"camera T_world_from_eye must be a rigid transform\nT\n{T.T}\n(T*T_t - I).max()\n{err}\n")
            self.T_world_from_eye = T

    def __repr__(self) -> str:
        return f"{type(self).__name__}({self.width}x{self.height}, f={self.f} c={self.c}"

    def uv_to_window_matrix(self) -> np.ndarray:
        return np.array([[self.f[0], 0, self.c[0]], [0, self.f[1], self.c[1]], [0, 0, 1]])

    # Host versions of the point maps, for tests and small inputs; batches of pixels go through
    # misc.warp_crops / misc.warp_image on the GPU.
    def world_to_eye(self, v: np.ndarray) -> np.ndarray:
        return (np.asarray(v) - self.T_world_from_eye[:3, 3]) @ self.T_world_from_eye[:3, :3]

    def eye_to_world(self, v: np.ndarray) -> np.ndarray:
        return np.asarray(v) @ self.T_world_from_eye[:3, :3].T + self.T_world_from_eye[:3, 3]

    def eye_to_window(self, v: np.ndarray) -> np.ndarray:
        v = np.asarray(v)
        return v[..., :2] / v[..., 2, None] * self.f + self.c

    def window_to_eye(self, w: np.ndarray) -> np.ndarray:
        q = (np.asarray(w) - self.c) / self.f
        ray = np.concatenate([q, np.ones_like(q[..., :1])], axis=-1)
        return ray / np.linalg.norm(ray, axis=-1, keepdims=True)

    def world_to_window(self, v: np.ndarray) -> np.ndarray:
        return self.eye_to_window(self.world_to_eye(v))


class PinholePlaneCameraModel(CameraModel):
    """The only camera model on the FoundPose inference path (reference :672-680)."""
