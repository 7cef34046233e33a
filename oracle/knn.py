"""Oracle: brute-force k-NN with faiss `IndexFlatL2` / `IndexFlatIP` semantics.  TEST INFRASTRUCTURE ONLY.

Follows what utils/knn_util.py:38-106 asks of faiss 1.8.0 (a third-party dependency pinned in
conda_foundpose_gpu.yaml:19, absent from /root/reference and from this image):
  * metric "l2":     squared L2 distances, ascending, float32; indices int64
                     (utils/template_util.py:26 "The distances returned by faiss are squared")
  * metric "cosine": rows L2-normalised at fit and search time, inner product, distance = 1 - sim
                     (utils/knn_util.py:54-60, 93-98)
faiss's published algorithm for flat indexes (faiss/utils/distances.cpp, v1.8.0):
  * nq >= distance_compute_blas_threshold (= 20): exhaustive_L2sqr_blas - queries in blocks of 4096, database in
    blocks of 1024, d(x, y) = ||x||^2 + ||y||^2 - 2 <x, y> with the inner products from one sgemm per block pair,
    evaluated in fp32 with negatives clamped to 0, candidates pushed into a per-query heap -> `knn_l2`,
    `knn_l2_blocked` (same arithmetic, database streamed block by block for banks that do not fit a dense
    distance matrix);
  * nq < 20: exhaustive_L2sqr_seq - every distance evaluated directly as sum_i (x_i - y_i)^2 (fvec_L2sqr), no
    norm expansion, so the low-order bits differ from the blas branch -> `knn_l2_direct`.
`knn_l2_faiss` dispatches on nq like faiss does.  Ties are broken by ascending index (canonical rule,
SURVEY.md §8c(i)); faiss's heap order for exact ties is implementation-defined.
Parity unpinned against faiss binaries (faiss is absent from /root/reference and from this image).
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


def knn_l2_direct(query: torch.Tensor, bank: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """faiss exhaustive_L2sqr_seq (nq < 20): d = sum_i (q_i - x_i)^2 accumulated in fp32 (fvec_L2sqr)."""
    q = query.to(torch.float32)
    x = bank.to(torch.float32)
    d = torch.stack([((x - q[i]) ** 2).sum(dim=1) for i in range(q.shape[0])]) if q.shape[0] else \
        torch.empty(0, x.shape[0])
    return _topk_smallest(d, k)


FAISS_BLAS_THRESHOLD = 20   # faiss::distance_compute_blas_threshold


def knn_l2_faiss(query: torch.Tensor, bank: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """IndexFlatL2.search as faiss dispatches it: direct distances below 20 queries, the blas expansion above."""
    if query.shape[0] < FAISS_BLAS_THRESHOLD:
        return knn_l2_direct(query, bank, k)
    return knn_l2(query, bank, k)


def merge_topk(d_a: torch.Tensor, i_a: torch.Tensor, d_b: torch.Tensor, i_b: torch.Tensor, k: int
               ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Merges two per-query candidate lists by (distance, index) ascending."""
    d = torch.cat([d_a, d_b], dim=1)
    i = torch.cat([i_a, i_b], dim=1)
    # lexicographic: stable sort by index first, then stable sort by distance
    o1 = torch.sort(i, dim=1, stable=True).indices
    d, i = torch.gather(d, 1, o1), torch.gather(i, 1, o1)
    o2 = torch.sort(d, dim=1, stable=True).indices
    return torch.gather(d, 1, o2)[:, :k].contiguous(), torch.gather(i, 1, o2)[:, :k].contiguous()


def knn_l2_blocked(query: torch.Tensor, bank_blocks, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """exhaustive_L2sqr_blas over a bank delivered block by block.

    `bank_blocks` yields (first_row, rows fp32 [n, d]) in ascending row order; every block contributes its k best
    (same fp32 expansion as `knn_l2`), merged by (distance, global index) - what faiss's per-query heap holds after
    the last database block.  Used for the full-bank search K4 of BASELINE configs 3 and 5, where the bank
    (10.24 M rows) is far larger than a dense distance matrix allows.
    """
    best_d = best_i = None
    for row0, rows in bank_blocks:
        kk = min(k, rows.shape[0])
        d, i = knn_l2(query, rows, kk)
        i = i + int(row0)
        if best_d is None:
            best_d, best_i = d, i
        else:
            best_d, best_i = merge_topk(best_d, best_i, d, i, k)
    if best_d is not None and best_d.shape[1] < k:
        pad = k - best_d.shape[1]
        best_d = torch.cat([best_d, torch.full((best_d.shape[0], pad), float("inf"))], dim=1)
        best_i = torch.cat([best_i, torch.full((best_i.shape[0], pad), -1, dtype=torch.int64)], dim=1)
    return best_d, best_i


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
