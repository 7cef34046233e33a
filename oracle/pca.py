"""Oracle: PCA projection of descriptors.  TEST INFRASTRUCTURE ONLY.

utils/projector_util.py:66-69 calls sklearn.decomposition.PCA.transform on numpy arrays; with
`whiten=False` (always the case after projector_from_tensordict, :128, 139-140) scikit-learn
computes  X @ components_.T - mean_ @ components_.T  in fp32.  Restated with numpy.
"""

from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch


def pca_transform(x: torch.Tensor, components: torch.Tensor, mean: torch.Tensor) -> torch.Tensor:
    xn = x.detach().cpu().numpy().astype(np.float32)
    c = components.detach().cpu().numpy().astype(np.float32)
    m = mean.detach().cpu().numpy().astype(np.float32)
    out = xn @ c.T
    out -= m.reshape(1, -1) @ c.T
    return torch.from_numpy(out)


def project_features(feat_vectors: torch.Tensor, projector_dicts: List[Dict]) -> torch.Tensor:
    """utils/projector_util.py:71-88 over a list of `pca_projector` tensordicts."""
    for pd in projector_dicts:
        p = pd["pca_projector"]
        feat_vectors = pca_transform(feat_vectors, p["components"], p["mean"])
    return feat_vectors
