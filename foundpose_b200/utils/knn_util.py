"""Brute-force k-NN on the B200 - drop-in for the reference's utils/knn_util.py (faiss wrapper).

`KNN(k, metric).fit(data)` / `.search(data) -> (distances [N,k] fp32, indices [N,k] int64)` with the
semantics the reference obtains from faiss.IndexFlatL2 / IndexFlatIP (utils/knn_util.py:38-106):
metric "l2" returns SQUARED L2 distances in ascending order; metric "cosine" L2-normalises both sides
and returns 1 - cosine similarity.  Rows are stored in fp16 with fp32 ||x||^2 (north-star layout);
the search is `fp_knn_search_items` (tcgen05 distance tiles + register top-k).  Results are returned
on the device of the query tensor, as the reference does (:102-104).
"""

import math
import os
from typing import Any, Optional, Tuple

import torch

from foundpose_b200 import _native

_MAX_K = 16
_SPLIT_BELOW_ITEMS = 74     # fewer query blocks than half the SMs -> split the bank
_SPLIT_MIN_ROWS = 16384
_NUM_SMS = 148
_TARGET_ITEMS = int(os.environ.get("FOUNDPOSE_KNN_TARGET_ITEMS", "592"))   # upper bound on items of a split search


def _choose_num_chunks(n_q: int) -> int:
    """Bank chunks per query block for the split path: the fewest chunks whose n_q x chunks items fill whole
    waves of the 148 persistent CTAs (every extra item per SM costs a pipeline restart and a merge: one chunk per
    SM reached 57-88% of the HBM peak on 0.8-7.9 GB banks where four reached 45-75%)."""
    effs = []
    for c in range(1, max(1, _TARGET_ITEMS // n_q) + 1):
        waves = n_q * c / _NUM_SMS
        effs.append(waves / math.ceil(waves))
    best = max(effs)
    return 1 + next(i for i, e in enumerate(effs) if e >= best - 0.03)


def _device_for(t: torch.Tensor) -> torch.device:
    return t.device if t.is_cuda else torch.device("cuda", torch.cuda.current_device())


def _pad64(x: torch.Tensor) -> torch.Tensor:
    d = x.shape[1]
    dp = (d + 63) // 64 * 64
    x = x.to(torch.float32)
    if dp != d:
        x = torch.nn.functional.pad(x, (0, dp - d))
    return x.contiguous()


class KNN:
    """K nearest neighbor search."""

    def __init__(self, k: int = 1, metric: str = "l2", radius: Optional[float] = None,
                 res: Optional[Any] = None) -> None:
        self.index: Any = None
        self.k: int = k
        self.metric: str = metric
        self.radius: Optional[float] = radius
        self.res: Optional[Any] = res
        self._bank16: Optional[torch.Tensor] = None
        self._bank_sqnorm: Optional[torch.Tensor] = None

    @classmethod
    def from_packed(cls, bank16: torch.Tensor, bank_sqnorm: torch.Tensor, k: int = 1, metric: str = "l2") -> "KNN":
        """An index over already packed fp16 rows (a zero-copy view of an ObjectIndex segment)."""
        self = cls(k=k, metric=metric)
        self._bank16, self._bank_sqnorm = bank16, bank_sqnorm
        self.index = self
        return self

    def fit(self, data: torch.Tensor) -> None:
        """Creates index from provided vectors of shape (num_vectors, dimensionality)."""
        if self.metric not in ("l2", "cosine"):
            raise ValueError(f"Metric {self.metric} is not supported.")
        dev = _device_for(data)
        x = _pad64(data.detach().to(dev))
        self._bank16 = _native.convert_rows_f16(x, l2_normalize=(self.metric == "cosine"))
        self._bank_sqnorm = _native.row_sqnorm_f16(self._bank16)
        self.index = self

    def search(self, data: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Finds nearest neighbors; returns (distances, indices) of the k nearest neighbors."""
        if self.metric not in ("l2", "cosine"):
            raise ValueError(f"Metric {self.metric} is not supported.")
        if self.radius is not None:
            # The reference calls index.range_search_with_radius, which faiss does not provide
            # (dead code).
            raise NotImplementedError("radius search is dead code in the reference and is not provided")
        if self._bank16 is None:
            raise RuntimeError("KNN.fit must be called before KNN.search")
        if self.k > _MAX_K:
            raise NotImplementedError(f"k={self.k} > {_MAX_K} is not supported by the B200 k-NN kernel")
        out_device = data.device
        dev = self._bank16.device
        q = _pad64(data.detach().to(dev))
        nq = q.shape[0]
        dist = torch.empty((nq, self.k), dtype=torch.float32, device=dev)
        idx = torch.empty((nq, self.k), dtype=torch.int64, device=dev)
        if nq > 0:
            cosine = self.metric == "cosine"
            q16 = _native.convert_rows_f16(q, l2_normalize=cosine)
            qn = _native.row_sqnorm_f16(q16)
            n_q = _native.knn_num_items(nq)
            nb = self._bank16.shape[0]
            metric = 1 if cosine else 0
            # Few query blocks against a large bank: split the bank so that all SMs stream it.
            num_chunks = 1
            if n_q < _SPLIT_BELOW_ITEMS and nb >= _SPLIT_MIN_ROWS:
                want = _choose_num_chunks(n_q)
                chunk_rows = max(256, ((nb + want - 1) // want + 255) // 256 * 256)
                num_chunks = (nb + chunk_rows - 1) // chunk_rows
            if num_chunks > 1:
                q_pad = n_q * 128
                items = _native.new_knn_items(n_q * num_chunks, dev)
                _native.knn_items_split(items, nq, 0, nb, num_chunks, chunk_rows)
                part_d = torch.empty((num_chunks * q_pad, self.k), dtype=torch.float32, device=dev)
                part_i = torch.full((num_chunks * q_pad, self.k), -1, dtype=torch.int64, device=dev)
                _native.knn_search_items(q16, qn, self._bank16, self._bank_sqnorm, items, n_q * num_chunks,
                                         metric, self.k, part_d, part_i)
                _native.knn_merge(part_d, part_i, num_chunks, q_pad, nq, self.k, chunk_rows, nb, cosine, dist, idx)
            else:
                items = _native.new_knn_items(n_q, dev)
                _native.knn_items_dense(items, nq, 0, nb)
                _native.knn_search_items(q16, qn, self._bank16, self._bank_sqnorm, items, n_q, metric, self.k,
                                         dist, idx)
            if cosine:
                dist = 1.0 - dist  # cosine similarity -> cosine distance (reference :98)
        return dist.to(out_device), idx.to(out_device)

    def serialize_index(self) -> None:
        pass

    def deserialize_index(self) -> None:
        pass
