"""oracle/crop.py against the golden outputs of the reference's own crop stage (utils/misc.py) and of
cv2.remap (tests/golden/make_golden_crop.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from foundpose_b200 import synthetic
from oracle import crop as ocrop

GOLD_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_crop_v1.pt")


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD_PATH, weights_only=False)


def test_remap_restatement_is_bit_exact_with_opencv(gold):
    src, mx, my = gold["remap/src"].numpy(), gold["remap/map_x"].numpy(), gold["remap/map_y"].numpy()
    lin = ocrop.remap_linear_f32(src, mx, my)
    assert np.array_equal(lin, gold["remap/linear"].numpy())
    assert np.array_equal(lin, gold["remap/area"].numpy())          # remap treats INTER_AREA as INTER_LINEAR
    near = ocrop.remap_nearest(gold["remap/src_u8"].numpy(), mx, my)
    assert np.array_equal(near, gold["remap/nearest_u8"].numpy())


def test_crop_stage_matches_the_reference(gold):
    cases = synthetic.make_crop_cases()
    assert len(cases) == gold["num_cases"]
    for i, case in enumerate(cases):
        p = f"case{i}/"
        crop_box = ocrop.calc_crop_box(case["box"], make_square=True)
        assert np.allclose(crop_box, gold[p + "crop_box"].numpy(), rtol=0, atol=1e-12)
        out = ocrop.crop_instance(case["image"], case["mask"], case["box"], case["f"], case["c"],
                                  case["T_world_from_eye"], case["crop_size"], case["crop_rel_pad"])
        cam = out["camera"]
        assert np.allclose(np.array(cam["f"], dtype=np.float64), gold[p + "cam_f"].numpy(), rtol=1e-6)
        assert np.allclose(np.array(cam["c"], dtype=np.float64), gold[p + "cam_c"].numpy(), rtol=0, atol=0)
        assert np.allclose(cam["T_world_from_eye"], gold[p + "cam_T"].numpy(), rtol=0, atol=1e-12)
        # bit-exact pixels: same fixed-point source coordinates, same fp32 tap arithmetic
        assert np.array_equal(out["image"], gold[p + "image"].numpy()), i
        assert np.array_equal(out["mask"], gold[p + "mask"].numpy()), i
        assert np.array_equal(out["box"], gold[p + "box"].numpy()), i
