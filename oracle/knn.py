"""Oracle: brute-force k-NN with faiss `IndexFlatL2` / `IndexFlatIP` semantics.  TEST INFRASTRUCTURE ONLY.

Follows what utils/knn_util.py:38-106 asks of faiss 1.8.0 (a third-party dependency pinned in
conda_foundpose_gpu.yaml:19, absent from /root/reference and from this image):
  * metric "l2":     squared L2 distances, ascending, float32; indices int64
                     (utils/template_util.py:26 "The distances returned by faiss are squared")
  * metric "cosine": rows L2-normalised at fit and search time, inner product, distance = 1 - sim
                     (utils/knn_util.py:54-60, 93-98)
faiss's published algorithm for flat indexes with >= 20 queries (exhaustive_L2sqr_blas) is
d(x, y) = ||x||^2 + ||y||^2 - 2 <x, y> evaluated in fp32 with negatives clamped to 0; we restate
exactly that.  Ties are broken by ascending index (canonical rule).
"""

from __future__ import annotations

from typing import Optional, Tuple

import torch


def _topk_smallest(dist: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """k smallest per row, ascending, ties -> ascending column index."""
    if k <= 8 and dist.shape[1] > 4 * k:
        # k passes of argmin (which returns the FIRST minimal index) - same result as the stable
        # sort below, much cheaper for the k = 1 / 3 / 5 searches of the path.
        work = dist.clone()
        vals, idxs = [], []
        rows = torch.arange(dist.shape[0])
        for _ in range(k):
            i = torch.argmin(work, dim=1)
            vals.append(work[rows, i])
            idxs.append(i)
            work[rows, i] = float("inf")
        return torch.stack(vals, dim=1).contiguous(), torch.stack(idxs, dim=1).contiguous()
    order = torch.sort(dist, dim=1, stable=True)
    return order.values[:, :k].contiguous(), order.indices[:, :k].contiguous()


def l2_distances(query: torch.Tensor, bank: torch.Tensor) -> torch.Tensor:
    q = query.to(torch.float32)
    x = bank.to(torch.float32)
    qn = (q * q).sum(dim=1, keepdim=True)
    xn = (x * x).sum(dim=1).unsqueeze(0)
    d = qn + xn - 2.0 * (q @ x.t())
    return torch.clamp_min(d, 0.0)


def knn_l2(query: torch.Tensor, bank: torch.Tensor, k: int, chunk: int = 4096) -> Tuple[torch.Tensor, torch.Tensor]:
    """(squared distances [Nq,k] fp32 ascending, indices [Nq,k] int64)."""
    outs_d, outs_i = [], []
    for s in range(0, query.shape[0], chunk):
        d = l2_distances(query[s:s + chunk], bank)
        dd, ii = _topk_smallest(d, k)
        outs_d.append(dd)
        outs_i.append(ii)
    if not outs_d:
        return torch.empty(0, k), torch.empty(0, k, dtype=torch.int64)
    return torch.cat(outs_d), torch.cat(outs_i)


def knn_cosine(query: torch.Tensor, bank: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """(1 - cosine similarity [Nq,k] ascending, indices [Nq,k] int64)."""
    x = bank / torch.linalg.norm(bank, dim=1, keepdim=True)
    q = query / torch.linalg.norm(query, dim=1, keepdim=True)
    sim = q.to(torch.float32) @ x.to(torch.float32).t()
    dd, ii = _topk_smallest(-sim, k)
    return 1.0 - (-dd), ii


def topk_margin(query: torch.Tensor, bank: torch.Tensor, k: int) -> torch.Tensor:
    """Relative gap between every returned neighbour and the next candidate, computed in fp64.

    margin[i, j] = min(d_{j+1} - d_j, d_j - d_{j-1}) / max(d_j, eps): index j of query i is only
    required to match bit-exactly where this exceeds 1e-4.
    """
    q = query.to(torch.float64)
    x = bank.to(torch.float64)
    d = torch.cdist(q, x).square()
    kk = min(k + 1, bank.shape[0])
    vals = torch.sort(d, dim=1, stable=True).values[:, :kk]
    if kk <= k:
        vals = torch.cat([vals, torch.full((vals.shape[0], k + 1 - kk), float("inf"), dtype=vals.dtype)], dim=1)
    gap_next = vals[:, 1:k + 1] - vals[:, :k]
    gap_prev = torch.cat([torch.full_like(vals[:, :1], float("inf")), gap_next[:, :-1]], dim=1)
    gap = torch.minimum(gap_next, gap_prev)
    return (gap / vals[:, :k].clamp_min(1e-12)).to(torch.float32)


class KNN:
    """Drop-in for utils/knn_util.py:10-112 running the oracle arithmetic on the CPU."""

    def __init__(self, k: int = 1, metric: str = "l2", radius: Optional[float] = None, res=None) -> None:
        self.k = k
        self.metric = metric
        self.radius = radius
        self.res = res
        self.index: Optional[torch.Tensor] = None

    def fit(self, data: torch.Tensor) -> None:
        if self.metric not in ("l2", "cosine"):
            raise ValueError(f"Metric {self.metric} is not supported.")
        self.index = data.detach().cpu().to(torch.float32).contiguous()

    def search(self, data: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        if self.metric == "l2":
            return knn_l2(data.cpu(), self.index, self.k)
        elif self.metric == "cosine":
            return knn_cosine(data.cpu(), self.index, self.k)
        raise ValueError(f"Metric {self.metric} is not supported.")
