"""PCA projection of descriptors - mirror of the reference's utils/projector_util.py.

`PCAProjector.transform` evaluates X @ C^T - mean @ C^T (what sklearn's PCA.transform computes with
whiten=False, reference :66-69) as a tcgen05 GEMM with the bias -(mean @ C^T) fused into the
epilogue (`fp_pca_project`).  `fit` is offline (gen_repre) and stays on scikit-learn.
"""

from typing import Any, Dict, List, Optional

import numpy as np
import torch

from foundpose_b200 import _native
from foundpose_b200.utils import logging
from foundpose_b200.utils.misc import array_to_tensor, tensor_to_array

logger: logging.Logger = logging.get_logger()


class Projector:
    """An abstract class for a projector."""

    def fit(self, data_x: torch.Tensor, data_y: Optional[torch.Tensor] = None, **kwargs: Any) -> None:
        raise NotImplementedError

    def transform(self, data_x: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError


class _PCAState:
    """Minimal stand-in for a fitted sklearn PCA (attribute names follow scikit-learn)."""

    def __init__(self, n_components: int, whiten: bool = False) -> None:
        self.n_components = n_components
        self.whiten = whiten
        self.components_ = None
        self.mean_ = None
        self.explained_variance_ = None
        self.explained_variance_ratio_ = None
        self.singular_values_ = None
        self.noise_variance_ = None


class PCAProjector(Projector):
    def __init__(self, n_components: int, whiten: bool = False, **kwargs: Dict[str, Any]) -> None:
        self.n_components: int = n_components
        self.whiten: bool = whiten
        self.pca: Any = _PCAState(n_components=n_components, whiten=whiten)
        self._device_state: Dict[str, Dict[str, torch.Tensor]] = {}

    def fit(self, data_x: torch.Tensor, data_y: Optional[torch.Tensor] = None, **kwargs: Any) -> None:
        """PCA fit of the offline bank build (reference utils/projector_util.py:53-64, scripts/gen_repre.py:272-284).

        CUDA tensors are fitted on the GPU with the algorithm scikit-learn calls "covariance_eigh"
        (sklearn/decomposition/_pca.py; its `svd_solver="auto"` takes it for tall data with few features, but picks the
        non-deterministic "randomized" solver for 1024-d features with 256 components, so a GPU-built bank is comparable
        to the reference's only up to the subspace, not bit for bit): covariance of the centred samples,
        symmetric eigendecomposition, eigenvectors flipped so that the largest entry of every component is positive
        (`svd_flip(u_based_decision=False)`).  The covariance - the O(n D^2) part - runs on the tcgen05 GEMM with the
        fp32 samples split into two fp16 terms (x = hi + lo, the lo.lo product is below fp32 resolution); the D x D
        eigendecomposition is cuSOLVER's through torch.linalg.eigh (float64).  CPU tensors go to scikit-learn exactly
        as the reference does.
        """
        if "max_samples" in kwargs:
            if data_x.shape[0] > kwargs["max_samples"]:
                perm = torch.randperm(data_x.shape[0])
                data_x = data_x[perm[: kwargs["max_samples"]].to(data_x.device)]
        self._device_state = {}
        if not data_x.is_cuda:
            from sklearn.decomposition import PCA

            self.pca = PCA(n_components=self.n_components, whiten=self.whiten)
            self.pca.fit(tensor_to_array(data_x))
            return
        x = data_x.detach().to(torch.float32)
        n, d = x.shape
        if not 1 <= self.n_components <= min(n, d):
            raise ValueError(f"n_components={self.n_components} must be between 1 and min(n_samples, n_features)="
                             f"{min(n, d)}")
        mean = x.mean(dim=0)
        d_pad = (d + 127) // 128 * 128
        n_pad = (n + 63) // 64 * 64
        xt = torch.zeros((d_pad, n_pad), dtype=torch.float32, device=x.device)      # [features, samples]
        xt[:d, :n] = (x - mean).t()
        hi = xt.to(torch.float16)
        lo = (xt - hi.to(torch.float32)).to(torch.float16)
        zero_bias = torch.zeros(d_pad, dtype=torch.float32, device=x.device)
        g_hh = torch.empty((d_pad, d_pad), dtype=torch.float32, device=x.device)
        g_hl = torch.empty((d_pad, d_pad), dtype=torch.float32, device=x.device)
        _native.gemm_tn_f16(hi, hi, _native.EPI_BIAS_F32, bias=zero_bias, out_f32=g_hh)
        _native.gemm_tn_f16(hi, lo, _native.EPI_BIAS_F32, bias=zero_bias, out_f32=g_hl)
        cov = (g_hh + g_hl + g_hl.t())[:d, :d].to(torch.float64)
        cov = 0.5 * (cov + cov.t()) / (n - 1)
        eigenvals, eigenvecs = torch.linalg.eigh(cov)
        eigenvals = torch.flip(eigenvals, dims=(0,)).clamp_min(0.0)
        vt = torch.flip(eigenvecs, dims=(1,)).t()                                   # rows = components
        # svd_flip(u_based_decision=False): the entry of largest magnitude of every row becomes positive
        idx = vt.abs().argmax(dim=1)
        signs = torch.sign(vt[torch.arange(vt.shape[0], device=vt.device), idx])
        signs[signs == 0] = 1.0
        vt = vt * signs[:, None]
        k = self.n_components
        state = _PCAState(n_components=k, whiten=self.whiten)
        state.components_ = vt[:k].to(torch.float32).cpu().numpy()
        state.mean_ = mean.cpu().numpy()
        state.explained_variance_ = eigenvals[:k].to(torch.float32).cpu().numpy()
        state.explained_variance_ratio_ = (eigenvals[:k] / eigenvals.sum()).to(torch.float32).cpu().numpy()
        state.singular_values_ = torch.sqrt(eigenvals[:k] * (n - 1)).to(torch.float32).cpu().numpy()
        state.noise_variance_ = float(eigenvals[k:].mean()) if k < min(n, d) else 0.0
        state.n_samples_, state.n_features_in_ = n, d
        self.pca = state

    def device_state(self, device: torch.device) -> Dict[str, torch.Tensor]:
        """fp16 components (rows padded to a multiple of 128) and the fused bias on `device`."""
        key = str(device)
        if key not in self._device_state:
            if bool(np.asarray(getattr(self.pca, "whiten", False)).any()):
                raise NotImplementedError("whitened PCA never occurs after projector_from_tensordict")
            comp = torch.as_tensor(np.asarray(self.pca.components_), dtype=torch.float32)
            mean = torch.as_tensor(np.asarray(self.pca.mean_), dtype=torch.float32)
            d_out, d_in = comp.shape
            assert d_in % 64 == 0, f"input feature dim {d_in} must be a multiple of 64"
            d_pad = (d_out + 127) // 128 * 128
            comp_p = torch.zeros(d_pad, d_in)
            comp_p[:d_out] = comp
            # The GEMM multiplies by the fp16-rounded components, so the folded bias -(mean . C^T) uses the same
            # rounded values: out = (x - mean) . C16^T exactly as if the mean had been subtracted first.
            comp16 = comp.to(torch.float16).to(torch.float32)
            bias = torch.zeros(d_pad)
            bias[:d_out] = torch.from_numpy(-(mean.reshape(1, -1).numpy() @ comp16.numpy().T).reshape(-1))
            self._device_state[key] = {
                "components16": comp_p.to(device, torch.float16).contiguous(),
                "bias": bias.to(device).contiguous(),
                "d_out": d_out,
            }
        return self._device_state[key]

    def transform(self, data_x: torch.Tensor) -> torch.Tensor:
        if not data_x.is_cuda:
            raise ValueError("foundpose_b200 runs on CUDA tensors only (no CPU fallback)")
        st = self.device_state(data_x.device)
        x16 = _native.convert_rows_f16(data_x.to(torch.float32).contiguous())
        m = x16.shape[0]
        d_pad = st["components16"].shape[0]
        out = torch.empty((m, d_pad), dtype=torch.float32, device=data_x.device)
        if m > 0:
            _native.pca_project(x16, st["components16"], st["bias"], out, None)
        return out[:, : st["d_out"]]


def compose_projectors(projectors: List[Projector]) -> Projector:
    """One PCA projector equal to applying `projectors` in turn (project_features, reference :71-88).

    transform_i(x) = (x - m_i) C_i^T is affine, so the chain collapses to x A^T + b with A = C_n ... C_1 (float64):
    the batched pipeline runs ONE GEMM whatever the number of projectors stored in repre.pth.  Returned as a
    PCAProjector with components_ = A and mean_ chosen so that -mean . A^T = b (least squares; exact whenever b lies
    in the row space of A, which holds for a chain of PCA projections)."""
    if len(projectors) == 1:
        return projectors[0]
    a, b = None, None
    for proj in projectors:
        if not isinstance(proj, PCAProjector):
            raise ValueError(f"Unknown projector type: {type(proj)}")
        c = np.asarray(proj.pca.components_, dtype=np.float64)
        m = np.asarray(proj.pca.mean_, dtype=np.float64)
        if a is None:
            a, b = c, -(m @ c.T)
        else:
            a, b = c @ a, (b - m) @ c.T
    out = PCAProjector(n_components=a.shape[0], whiten=False)
    state = _PCAState(n_components=a.shape[0])
    state.components_ = a.astype(np.float32)
    state.mean_ = (-np.linalg.lstsq(a, b, rcond=None)[0]).astype(np.float32)
    out.pca = state
    return out


def project_features(feat_vectors: torch.Tensor, projectors: List[Projector], batch_size: int = 4096) -> torch.Tensor:
    """Projects (num_features, feat_dims) feature vectors with every projector in turn."""
    for projector in projectors:
        feat_vectors = projector.transform(feat_vectors)
    return feat_vectors


def projector_to_tensordict(projector: Projector) -> Dict[str, Any]:
    """Converts a feature projector to a tensordict (same keys as the reference :91-113)."""
    if isinstance(projector, PCAProjector):
        return {
            "pca_projector": {
                "components": torch.tensor(np.asarray(projector.pca.components_)),
                "explained_variance": torch.tensor(np.asarray(projector.pca.explained_variance_)),
                "explained_variance_ratio": torch.tensor(np.asarray(projector.pca.explained_variance_ratio_)),
                "singular_values": torch.tensor(np.asarray(projector.pca.singular_values_)),
                "mean": torch.tensor(np.asarray(projector.pca.mean_)),
                "noise_variance": torch.tensor(np.asarray(projector.pca.noise_variance_)),
                "whiten": torch.tensor(projector.whiten),
            }
        }
    else:
        raise ValueError(f"Unknown projector type: {type(projector)}")


def projector_from_tensordict(projector_dict: Dict[str, Any]) -> Projector:
    """Builds a projector from a tensordict (reference :116-145)."""
    if "pca_projector" in projector_dict:
        p = projector_dict["pca_projector"]
        pca = _PCAState(n_components=len(p["components"]))
        pca.components_ = tensor_to_array(p["components"])
        pca.explained_variance_ = tensor_to_array(p["explained_variance"])
        pca.explained_variance_ratio_ = tensor_to_array(p["explained_variance_ratio"])
        pca.singular_values_ = tensor_to_array(p["singular_values"])
        pca.mean_ = tensor_to_array(p["mean"])
        pca.noise_variance_ = tensor_to_array(p["noise_variance"])
        pca.whiten_ = tensor_to_array(p["whiten"])
        # As in the reference, `whiten` stays False after loading (PCA(...) is rebuilt with defaults).
        projector = PCAProjector(n_components=len(pca.components_), whiten=False)
        projector.whiten = pca.whiten_
        projector.pca = pca
        return projector
    else:
        raise ValueError("Unknown projector type.")
