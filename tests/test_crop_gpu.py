"""GPU parity of the crop stage (csrc/crop_warp.cu) against the golden outputs of the
reference's utils/misc.py + cv2.remap and against oracle/crop.py.

Bar: bit-exact pixels, masks and boxes on the golden cases (same fixed-point source coordinates, same fp32
tap arithmetic).  On the full-size random case the source coordinates come from two different fp64
evaluation orders (numpy's BLAS vs the kernel), which may differ in the last bit exactly at a tie of the
1/32-pixel quantiser: at least 99.99% of pixels must be bit-identical and none may differ by more than 1/32
of the local contrast (stated in the test).
"""
import os

import numpy as np
import pytest
import torch

from foundpose_b200 import synthetic

pytestmark = pytest.mark.gpu
GOLD_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_crop_v1.pt")


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD_PATH, weights_only=False)


def _camera(case):
    from foundpose_b200.utils import structs
    h, w = case["image"].shape[:2]
    return structs.PinholePlaneCameraModel(w, h, case["f"], case["c"], case["T_world_from_eye"])


def test_crop_stage_vs_reference_golden(gold):
    from foundpose_b200.utils import misc
    for i, case in enumerate(synthetic.make_crop_cases()):
        p = f"case{i}/"
        crops, masks, boxes, cams = misc.crop_instances(
            torch.from_numpy(case["image"]).cuda(), torch.from_numpy(case["mask"]).cuda()[None],
            case["box"][None], _camera(case), case["crop_size"], case["crop_rel_pad"])
        assert crops.shape == (1, 3, case["crop_size"][1], case["crop_size"][0])
        hwc = crops[0].permute(1, 2, 0).cpu()
        assert torch.equal(hwc, gold[p + "image"]), (i, (hwc - gold[p + "image"]).abs().max())
        assert torch.equal(masks[0].cpu(), gold[p + "mask"]), i
        assert torch.equal(boxes[0].cpu(), gold[p + "box"]), i
        assert np.allclose(np.array(cams[0].f, dtype=np.float64), gold[p + "cam_f"].numpy(), rtol=1e-12)


def test_warp_image_signature_vs_golden(gold):
    """The reference-shaped single-image call (numpy in, numpy out) and batching over instances agree."""
    from foundpose_b200.utils import misc, structs
    cases = synthetic.make_crop_cases()
    case = cases[1]
    cam = _camera(case)
    crop_cam = misc.construct_crop_camera(misc.calc_crop_box(structs.AlignedBox2f(*case["box"]), make_square=True),
                                          cam, case["crop_size"], case["crop_rel_pad"])
    image = case["image"].astype(np.float32) / 255.0
    warped = misc.warp_image(cam, crop_cam, image, interpolation=misc.INTER_AREA)
    assert isinstance(warped, np.ndarray) and np.array_equal(warped, gold["case1/image"].numpy())
    wmask = misc.warp_image(cam, crop_cam, case["mask"], interpolation=misc.INTER_NEAREST)
    assert np.array_equal(wmask, gold["case1/mask"].numpy())
    gray = misc.warp_image(cam, crop_cam, torch.from_numpy(image[:, :, 1].copy()).cuda(), interpolation=misc.INTER_LINEAR)
    assert torch.equal(gray.cpu(), gold["case1/image"][:, :, 1])
    with pytest.raises(ValueError):
        misc.warp_image(cam, crop_cam, image, interpolation=2)


def test_full_size_batch_vs_oracle():
    """8 instances of a 480x640 image into 420x420 crops (BASELINE crop size) against oracle/crop.py."""
    from foundpose_b200.utils import misc, structs
    from oracle import crop as ocrop

    H, W, B = 480, 640, 8
    rng = np.random.RandomState(0)
    image = synthetic.make_scene_image(H, W, seed=1)
    T = synthetic._rigid_transform(rng)
    cam = structs.PinholePlaneCameraModel(W, H, (572.4, 573.6), (325.3, 242.0), T)
    masks, boxes = [], []
    for b in range(B):
        center = (rng.uniform(40, W - 40), rng.uniform(40, H - 40))
        radii = (rng.uniform(15, 150), rng.uniform(15, 120))
        m, bx = synthetic.make_instance_mask(H, W, center, radii, seed=10 + b)
        masks.append(m)
        boxes.append(bx)
    masks[3][:] = 0                                                     # an instance whose mask is empty
    crops, wmasks, new_boxes, cams = misc.crop_instances(
        torch.from_numpy(image).cuda(), torch.from_numpy(np.stack(masks)).cuda(), np.stack(boxes), cam, (420, 420), 0.2)
    torch.cuda.synchronize()
    assert new_boxes[3].tolist() == [0.0, 0.0, 0.0, 0.0] and int(wmasks[3].sum()) == 0
    for b in range(B):
        ref = ocrop.crop_instance(image, masks[b], boxes[b], cam.f, cam.c, T, (420, 420), 0.2)
        ours = crops[b].permute(1, 2, 0).cpu().numpy()
        same = (ours == ref["image"]).all(axis=-1)
        assert same.mean() >= 0.9999, (b, same.mean())
        assert np.abs(ours - ref["image"]).max() <= 1.0 / 32 + 1e-6     # a tie moves one tap weight by 1/32
        assert (wmasks[b].cpu().numpy() == ref["mask"]).mean() >= 0.9999
        assert np.abs(new_boxes[b].cpu().numpy() - ref["box"]).max() <= 1.0


def test_images_from_several_sources_and_fp32_input():
    """image_index selects the source image per instance; fp32 sources are sampled without the 1/255 scale."""
    from foundpose_b200.utils import misc, structs
    cases = synthetic.make_crop_cases()
    cams, crop_cams, imgs = [], [], []
    for case in (cases[0], cases[2]):
        cam = _camera(case)
        cams.append(cam)
        crop_cams.append(misc.construct_crop_camera(
            misc.calc_crop_box(structs.AlignedBox2f(*case["box"]), make_square=True), cam, (42, 42), 0.25))
        imgs.append(torch.from_numpy(case["image"]))
    stack = torch.stack(imgs).cuda()
    masks = torch.stack([torch.from_numpy(cases[0]["mask"]), torch.from_numpy(cases[2]["mask"])]).cuda()
    both, _, _ = misc.warp_crops(stack, masks, cams, crop_cams, image_index=[0, 1])
    for k in range(2):
        single, _, _ = misc.warp_crops(stack[k:k + 1], masks[k:k + 1], cams[k:k + 1], crop_cams[k:k + 1])
        assert torch.equal(both[k], single[0])
        src_f32 = torch.from_numpy(imgs[k].numpy().astype(np.float32) / 255.0)[None].cuda()   # IEEE division, as infer.py:396
        as_f32, _, _ = misc.warp_crops(src_f32, None, cams[k:k + 1], crop_cams[k:k + 1])
        assert torch.equal(as_f32[0], single[0])
