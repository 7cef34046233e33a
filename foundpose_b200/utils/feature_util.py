"""Query-point grid, mask filter and feature sampling - mirror of the reference's utils/feature_util.py:18-131."""

from typing import Tuple

import torch

from foundpose_b200 import _native
from foundpose_b200.utils import dinov2_utils, logging

logger: logging.Logger = logging.get_logger()


def make_feature_extractor(model_name: str, **kwargs) -> torch.nn.Module:
    if model_name.startswith("dinov2_"):
        return dinov2_utils.DinoFeatureExtractor(model_name=model_name, **kwargs)
    else:
        raise NotImplementedError(model_name)


def generate_grid_points(grid_size: Tuple[int, int], cell_size: float = 1.0) -> torch.Tensor:
    """2D coordinates at the centers of the cells of a regular grid, (grid_width, grid_height) order.

    Index arithmetic only (no feature data), evaluated with the same torch calls as the reference
    (utils/feature_util.py:25-52) so the coordinates are bit-identical.
    """
    grid_cols = int(grid_size[0] / cell_size)
    grid_rows = int(grid_size[1] / cell_size)
    cell_half_size = cell_size / 2.0
    x = torch.linspace(cell_half_size, grid_size[0] - cell_half_size, grid_cols, dtype=torch.float)
    y = torch.linspace(cell_half_size, grid_size[1] - cell_half_size, grid_rows, dtype=torch.float)
    grid_x, grid_y = torch.meshgrid(x, y, indexing="xy")
    return torch.vstack((grid_x.flatten(), grid_y.flatten())).T


def filter_points_by_box(points: torch.Tensor, box: Tuple[float, float, float, float]):
    """Keeps only points strictly inside the 2D box (x1, y1, x2, y2); returns (points, mask)."""
    x1, y1, x2, y2 = box
    valid_mask = torch.logical_and(
        torch.logical_and(points[:, 0] > x1, points[:, 0] < x2),
        torch.logical_and(points[:, 1] > y1, points[:, 1] < y2),
    )
    return points[valid_mask], valid_mask


def filter_points_by_mask(points: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """Keeps only points inside the mask (reference utils/feature_util.py:75-97), order preserved."""
    if not points.is_cuda:
        raise ValueError("foundpose_b200 runs on CUDA tensors only (no CPU fallback)")
    dev = points.device
    pts = points.to(torch.float32).contiguous()
    n = pts.shape[0]
    masks = mask.to(dev).ne(0).to(torch.uint8).reshape(1, mask.shape[0], mask.shape[1]).contiguous()
    out_points = torch.empty((1, max(n, 1), 2), dtype=torch.float32, device=dev)
    out_ids = torch.empty((1, max(n, 1)), dtype=torch.int32, device=dev)
    out_counts = torch.zeros((1,), dtype=torch.int32, device=dev)
    if n > 0:
        _native.filter_points_by_mask(pts, masks, out_points, out_ids, out_counts)
    count = int(out_counts.item())  # sizes the result, like boolean indexing does in the reference
    return out_points[0, :count]


def sample_feature_map_at_points(feature_map_chw: torch.Tensor, points: torch.Tensor,
                                 image_size: Tuple[int, int]) -> torch.Tensor:
    """Bilinear sampling of a (C, H, W) feature map at (N, 2) image points -> (N, C).

    Reference: utils/feature_util.py:100-131 (grid_sample, align_corners=False, zero padding).
    The extractor returns feature_maps as a permuted view of token-major data, so the
    `permute(1, 2, 0)` below is a zero-copy view in the normal flow.
    """
    if not feature_map_chw.is_cuda:
        raise ValueError("foundpose_b200 runs on CUDA tensors only (no CPU fallback)")
    c, h, w = feature_map_chw.shape
    tokens = feature_map_chw.permute(1, 2, 0).to(torch.float32).contiguous().reshape(1, h * w, c)
    pts = points.to(feature_map_chw.device, torch.float32).contiguous().reshape(1, -1, 2)
    n = pts.shape[1]
    out = torch.empty((n, c), dtype=torch.float32, device=feature_map_chw.device)
    if n > 0:
        if c % 4 != 0:
            raise ValueError("feature dimension must be a multiple of 4")
        _native.sample_features(tokens, h, w, pts, None, float(image_size[0]), float(image_size[1]), out, None)
    return out


def lift_2d_points_to_3d(points: torch.Tensor, depth_image: torch.Tensor, camera_model) -> torch.Tensor:
    """3D camera-space points of 2D image points with a depth image (reference utils/feature_util.py:132-155).

    Index arithmetic and a gather on O(1000) points, evaluated with the same torch calls as the reference.
    """
    device = points.device
    focal = 0.5 * (camera_model.f[0] + camera_model.f[1])    # the reference uses the average of fx and fy
    points_3d_in_cam = torch.hstack([
        points - torch.as_tensor(camera_model.c).to(torch.float32).to(device),
        focal * torch.ones(points.shape[0], 1).to(torch.float32).to(device),
    ])
    depths = depth_image[torch.floor(points[:, 1]).to(torch.int64), torch.floor(points[:, 0]).to(torch.int64)].reshape(-1, 1)
    points_3d_in_cam = points_3d_in_cam * (depths / points_3d_in_cam[:, 2].reshape(-1, 1))
    return points_3d_in_cam


def erode_mask_5x5(object_mask: torch.Tensor) -> torch.Tensor:
    """`kornia.morphology.erosion(mask, ones(5, 5))` (reference :181-188): minimum over the 5x5 window,
    pixels outside the image do not erode (kornia's default geodesic border)."""
    m = object_mask.reshape(1, 1, *object_mask.shape).to(torch.float32)
    eroded = -torch.nn.functional.max_pool2d(-m, kernel_size=5, stride=1, padding=2)
    return eroded.reshape(object_mask.shape).to(object_mask.dtype)


def get_visual_features_registered_in_3d(image_chw: torch.Tensor, depth_image_hw: torch.Tensor,
                                         object_mask: torch.Tensor, camera, T_model_from_camera: torch.Tensor,
                                         extractor: torch.nn.Module, grid_cell_size: float, debug: bool = False,
                                         feature_map_chw: torch.Tensor = None
                                         ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Features of one template registered in 3D (reference utils/feature_util.py:158-241), same outputs:
    (feat_vectors [n, C], vertex_ids [n] int32, vertices_in_model [n, 3]).

    `feature_map_chw` lets the offline bank build pass the map of this template out of a BATCHED
    extractor call (scripts/gen_repre.py runs the extractor once per template).
    """
    device = image_chw.device
    grid_points = generate_grid_points(grid_size=(image_chw.shape[2], image_chw.shape[1]),
                                       cell_size=grid_cell_size).to(device)
    object_mask_eroded = erode_mask_5x5(object_mask)
    query_points = filter_points_by_mask(grid_points, object_mask_eroded)
    vertices_in_cam = lift_2d_points_to_3d(points=query_points, depth_image=depth_image_hw, camera_model=camera)
    T = T_model_from_camera.to(device, torch.float32)
    vertices_in_model = vertices_in_cam @ T[:3, :3].T + T[:3, 3]          # geometry.transform_3d_points_torch
    vertex_ids = torch.arange(vertices_in_model.shape[0], dtype=torch.int32)
    if feature_map_chw is None:
        feature_map_chw = extractor(image_chw.unsqueeze(0))["feature_maps"][0]
    feat_vectors = sample_feature_map_at_points(feature_map_chw=feature_map_chw, points=query_points,
                                                image_size=(image_chw.shape[-1], image_chw.shape[-2])).detach()
    return feat_vectors, vertex_ids, vertices_in_model
