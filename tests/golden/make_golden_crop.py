"""Generates tests/golden/golden_crop_v1.pt: outputs of the REFERENCE'S OWN crop stage and of cv2.remap.

Run in the build container only (needs /root/reference and OpenCV, neither is used at test time):

    python tests/golden/make_golden_crop.py

What runs unmodified from /root/reference: utils/misc.py (calc_crop_box, construct_crop_camera,
warp_image), utils/structs.py (AlignedBox2f, PinholePlaneCameraModel), utils/geometry.py.  `cv2.remap`
is OpenCV's (4.13 in this image).  Each case follows scripts/infer.py:396-459.
"""

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from foundpose_b200 import synthetic  # noqa: E402


def main() -> None:
    import cv2
    from utils import misc as ref_misc
    from utils.structs import AlignedBox2f, PinholePlaneCameraModel

    out = {"cv2_version": cv2.__version__}
    cases = synthetic.make_crop_cases()
    out["num_cases"] = len(cases)
    for i, case in enumerate(cases):
        cam = PinholePlaneCameraModel(width=case["image"].shape[1], height=case["image"].shape[0], f=case["f"],
                                      c=case["c"], T_world_from_eye=case["T_world_from_eye"])
        image = case["image"].astype(np.float32) / 255.0                                   # infer.py:396
        box = AlignedBox2f(*[float(v) for v in case["box"]])
        crop_box = ref_misc.calc_crop_box(box=box, make_square=True)                       # infer.py:419
        crop_cam = ref_misc.construct_crop_camera(box=crop_box, camera_model_c2w=cam,
                                                  viewport_size=case["crop_size"],
                                                  viewport_rel_pad=case["crop_rel_pad"])   # infer.py:425
        interpolation = cv2.INTER_AREA if crop_box.width >= crop_cam.width else cv2.INTER_LINEAR
        warped = ref_misc.warp_image(cam, crop_cam, image, interpolation=interpolation)    # infer.py:437
        warped_mask = ref_misc.warp_image(cam, crop_cam, case["mask"], interpolation=cv2.INTER_NEAREST)
        ys, xs = warped_mask.nonzero()
        new_box = np.array(ref_misc.calc_2d_box(xs, ys))                                   # infer.py:449-450
        p = f"case{i}/"
        out[p + "crop_box"] = torch.tensor([crop_box.left, crop_box.top, crop_box.right, crop_box.bottom],
                                           dtype=torch.float64)
        out[p + "cam_f"] = torch.tensor([float(v) for v in crop_cam.f], dtype=torch.float64)
        out[p + "cam_c"] = torch.tensor([float(v) for v in crop_cam.c], dtype=torch.float64)
        out[p + "cam_T"] = torch.from_numpy(np.asarray(crop_cam.T_world_from_eye, dtype=np.float64).copy())
        out[p + "used_inter_area"] = bool(interpolation == cv2.INTER_AREA)
        out[p + "image"] = torch.from_numpy(warped.copy())
        out[p + "mask"] = torch.from_numpy(warped_mask.copy())
        out[p + "box"] = torch.from_numpy(np.asarray(new_box, dtype=np.float32).copy())

    # cv2.remap pinned directly on adversarial maps: half-way fixed-point ties, borders, far outside.
    rng = np.random.RandomState(7)
    src = rng.rand(37, 53, 3).astype(np.float32)
    src_u8 = (rng.rand(37, 53) > 0.5).astype(np.uint8) * 255
    mx = (rng.rand(64, 80) * 60 - 4).astype(np.float32)
    my = (rng.rand(64, 80) * 44 - 4).astype(np.float32)
    mx[:8] = np.round(mx[:8] * 64) / 64          # multiples of 1/64: exact ties of the 1/32 quantiser
    my[:8] = np.round(my[:8] * 64) / 64
    mx[8:12] = np.round(mx[8:12])                # integer coordinates
    my[8:12] = np.round(my[8:12]) + 0.5          # rounding ties of INTER_NEAREST
    mx[12, :4] = [-1.0, 52.0, 52.5, 1e6]
    my[12, :4] = [-1.0, 36.0, 36.5, -1e6]
    out["remap/src"] = torch.from_numpy(src)
    out["remap/src_u8"] = torch.from_numpy(src_u8)
    out["remap/map_x"] = torch.from_numpy(mx)
    out["remap/map_y"] = torch.from_numpy(my)
    out["remap/linear"] = torch.from_numpy(cv2.remap(src, mx, my, cv2.INTER_LINEAR))
    out["remap/area"] = torch.from_numpy(cv2.remap(src, mx, my, cv2.INTER_AREA))
    out["remap/nearest_u8"] = torch.from_numpy(cv2.remap(src_u8, mx, my, cv2.INTER_NEAREST))

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_crop_v1.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
