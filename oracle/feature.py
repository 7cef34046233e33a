"""Oracle: query-point grid, mask filter and feature sampling.  TEST INFRASTRUCTURE ONLY.

Restates utils/feature_util.py:25-131 with the same torch ops.
"""

from __future__ import annotations

from typing import Tuple

import torch


def generate_grid_points(grid_size: Tuple[int, int], cell_size: float = 1.0) -> torch.Tensor:
    """utils/feature_util.py:25-52."""
    grid_cols = int(grid_size[0] / cell_size)
    grid_rows = int(grid_size[1] / cell_size)
    half = cell_size / 2.0
    x = torch.linspace(half, grid_size[0] - half, grid_cols, dtype=torch.float)
    y = torch.linspace(half, grid_size[1] - half, grid_rows, dtype=torch.float)
    grid_x, grid_y = torch.meshgrid(x, y, indexing="xy")
    return torch.vstack((grid_x.flatten(), grid_y.flatten())).T


def filter_points_by_box(points: torch.Tensor, box) -> Tuple[torch.Tensor, torch.Tensor]:
    """utils/feature_util.py:55-72 (strict inequalities)."""
    x1, y1, x2, y2 = box
    valid = torch.logical_and(
        torch.logical_and(points[:, 0] > x1, points[:, 0] < x2),
        torch.logical_and(points[:, 1] > y1, points[:, 1] < y2),
    )
    return points[valid], valid


def filter_points_by_mask(points: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """utils/feature_util.py:75-97."""
    points_int = (points + 0.5).int()
    points_int, valid = filter_points_by_box(points_int, (0, 0, mask.shape[1], mask.shape[0]))
    return points[valid][mask[points_int[:, 1], points_int[:, 0]].bool()]


def sample_feature_map_at_points(feature_map_chw: torch.Tensor, points: torch.Tensor,
                                 image_size: Tuple[int, int]) -> torch.Tensor:
    """utils/feature_util.py:100-131: bilinear grid_sample, zero padding, align_corners=False."""
    uv = torch.div(2.0, torch.as_tensor(image_size)).to(points.device) * points - 1.0
    coords = uv.unsqueeze(0).unsqueeze(2)
    feats = torch.nn.functional.grid_sample(feature_map_chw.unsqueeze(0), coords, align_corners=False)
    return feats[0, :, :, 0].permute(1, 0)
