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
