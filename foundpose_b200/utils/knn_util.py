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
import weakref
from typing import Any, Optional, Tuple

import torch

from foundpose_b200 import _native

_MAX_K = 16
_SPLIT_BELOW_ITEMS = 74     # fewer query blocks than half the SMs -> split the bank
_SPLIT_MIN_ROWS = 16384
_PAIR_ROWS = 256            # query rows per item of the pair kernel (fp_knn_search_pair_items)
_PAIR_MIN_BANK_ROWS = 8192  # below this the per-item pipeline fill dominates: keep the 1-CTA kernel
# bank tiles (256 rows) between two sweep barriers of the pair kernel: 64 tiles = 12.6 MB of a 384-d bank, a small
# fraction of the L2, ~180 us of work per barrier
_SWEEP_SYNC_TILES = int(os.environ.get("FOUNDPOSE_KNN_SWEEP_SYNC_TILES", "64"))
_sm_count = [0]


def _num_sms() -> int:
    """SM count of the current device, read once through the library (cudaDevAttrMultiProcessorCount)."""
    if _sm_count[0] == 0:
        _sm_count[0] = _native.num_sms() if torch.cuda.is_available() else 148
    return _sm_count[0]
_TARGET_ITEMS = int(os.environ.get("FOUNDPOSE_KNN_TARGET_ITEMS", "592"))   # upper bound on items of a split search


def plan_pair_items(nq: int, nb: int, num_clusters: int) -> Tuple[list, list, int, int]:
    """Work items of a tensor-bound search (nq queries x nb bank rows) for the pair kernel.

    One item = 256 query rows sweeping a bank segment on one CTA pair.  Items that fill whole waves of the
    `num_clusters` persistent clusters sweep the whole bank and write final results; the remaining `rem` items of the
    last, partial wave are split into `chunks = num_clusters // rem` bank slices each, so that the tail wave is
    `1 / chunks` as long and every cluster has work in it.  Their partial top-k lists are merged by fp_knn_merge.

    Returns (direct_items, split_items, chunks, chunk_rows): item tuples (q_row0, q_rows, b_row0, b_rows, out_row0);
    split items write to a partial buffer laid out [chunks, rem * 256, k] (out_row0 relative to it).
    """
    n_pairs = (nq + _PAIR_ROWS - 1) // _PAIR_ROWS
    full = (n_pairs // num_clusters) * num_clusters
    rem = n_pairs - full
    chunks = num_clusters // rem if rem > 0 else 1
    chunk_rows = 0
    if chunks >= 2:
        chunk_rows = ((nb + chunks - 1) // chunks + 255) // 256 * 256
        chunks = (nb + chunk_rows - 1) // chunk_rows
    if chunks < 2:
        full, rem, chunks = n_pairs, 0, 1
    direct = [(i * _PAIR_ROWS, min(_PAIR_ROWS, nq - i * _PAIR_ROWS), 0, nb, i * _PAIR_ROWS) for i in range(full)]
    split = []
    q_pad = rem * _PAIR_ROWS
    for c in range(chunks if rem else 0):
        for j in range(rem):
            q0 = (full + j) * _PAIR_ROWS
            split.append((q0, min(_PAIR_ROWS, nq - q0), c * chunk_rows, max(0, min(chunk_rows, nb - c * chunk_rows)),
                          c * q_pad + j * _PAIR_ROWS))
    return direct, split, chunks, chunk_rows


def _choose_num_chunks(n_q: int) -> int:
    """Bank chunks per query block for the split path: the fewest chunks whose n_q x chunks items fill whole
    waves of the 148 persistent CTAs (every extra item per SM costs a pipeline restart and a merge: one chunk per
    SM reached 57-88% of the HBM peak on 0.8-7.9 GB banks where four reached 45-75%)."""
    effs = []
    for c in range(1, max(1, _TARGET_ITEMS // n_q) + 1):
        waves = n_q * c / _num_sms()
        effs.append(waves / math.ceil(waves))
    best = max(effs)
    return 1 + next(i for i, e in enumerate(effs) if e >= best - 0.03)


def _device_for(t: torch.Tensor) -> torch.device:
    return t.device if t.is_cuda else torch.device("cuda", torch.cuda.current_device())


def _pad64(x: torch.Tensor, width: Optional[int] = None) -> torch.Tensor:
    d = x.shape[1]
    dp = (d + 63) // 64 * 64 if width is None else width
    if dp < d:
        raise ValueError(f"query dimension {d} exceeds the index dimension {dp}")
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
        self.index = weakref.proxy(self)     # `knn.index` as in the reference, without a reference cycle
        return self

    def fit(self, data: torch.Tensor) -> None:
        """Creates index from provided vectors of shape (num_vectors, dimensionality)."""
        if self.metric not in ("l2", "cosine"):
            raise ValueError(f"Metric {self.metric} is not supported.")
        dev = _device_for(data)
        x = _pad64(data.detach().to(dev))
        self._bank16 = _native.convert_rows_f16(x, l2_normalize=(self.metric == "cosine"))
        self._bank_sqnorm = _native.row_sqnorm_f16(self._bank16)
        self.index = weakref.proxy(self)

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
        q = _pad64(data.detach().to(dev), self._bank16.shape[1])   # the index may hold wider, zero-padded rows
        nq = q.shape[0]
        dist = torch.empty((nq, self.k), dtype=torch.float32, device=dev)
        idx = torch.empty((nq, self.k), dtype=torch.int64, device=dev)
        if nq > 0:
            cosine = self.metric == "cosine"
            q16 = _native.convert_rows_f16(q, l2_normalize=cosine)
            qn = _native.row_sqnorm_f16(q16)
            dist, idx = self.search_packed(q16, qn, dist, idx)
            if dist._base is not None:
                # the pair path returns views of buffers it reuses for the next search of the same shape:
                # the public API hands out tensors the caller owns
                dist, idx = dist.clone(), idx.clone()
            if cosine:
                dist = 1.0 - dist  # cosine similarity -> cosine distance (reference :98)
        return dist.to(out_device), idx.to(out_device)

    def search_packed(self, q16: torch.Tensor, qn: torch.Tensor, dist: Optional[torch.Tensor] = None,
                      idx: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Search with already packed fp16 query rows (and their fp32 ||q||^2): no conversion, no host sync.

        Returns raw kernel outputs (squared L2 ascending, or inner products descending for metric "cosine").  On
        the pair path the returned tensors are views of buffers that the NEXT search of the same shape overwrites
        (no allocation in a steady-state loop): consume or copy them before searching again.
        Picks the pass structure from the problem shape: the pair kernel when the search is tensor-bound (>= 74
        query blocks x a large bank), bank slices over all SMs when it is HBM-bound (few query blocks x a large
        bank), one item per query block otherwise."""
        dev = q16.device
        nq, nb = q16.shape[0], self._bank16.shape[0]
        metric = 1 if self.metric == "cosine" else 0
        if dist is None:
            dist = torch.empty((nq, self.k), dtype=torch.float32, device=dev)
            idx = torch.empty((nq, self.k), dtype=torch.int64, device=dev)
        n_q = _native.knn_num_items(nq)
        if n_q >= _num_sms() // 2 and nb >= _PAIR_MIN_BANK_ROWS:
            return self._search_pairs(q16, qn, metric, dist, idx)
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
                _native.knn_merge(part_d, part_i, num_chunks, q_pad, nq, self.k, chunk_rows, nb, metric == 1,
                                  dist, idx)
                return dist, idx
        items = _native.new_knn_items(n_q, dev)
        _native.knn_items_dense(items, nq, 0, nb)
        _native.knn_search_items(q16, qn, self._bank16, self._bank_sqnorm, items, n_q, metric, self.k, dist, idx)
        return dist, idx

    def _search_pairs(self, q16: torch.Tensor, qn: torch.Tensor, metric: int, dist: torch.Tensor,
                      idx: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Tensor-bound regime: ONE launch of the pair kernel over whole-wave items (final results) and the split
        items of the tail wave (partial lists behind the final rows of the same buffers), then the tail merge."""
        dev = q16.device
        nq, nb = q16.shape[0], self._bank16.shape[0]
        key = (nq, nb, self.k, str(dev))
        plan = getattr(self, "_pair_plan", None)
        if plan is None or plan[0] != key:
            clusters = _num_sms() // 2
            direct, split, chunks, chunk_rows = plan_pair_items(nq, nb, clusters)
            # items i, i + C, ... run on cluster i % C: the items of one wave sweep the whole bank together and meet
            # at the kernel's sweep barrier (6th field = participants of the wave)
            direct = [it + (min(clusters, len(direct) - (i // clusters) * clusters),) for i, it in enumerate(direct)]
            split = [(q0, qr, b0, br, o0 + nq) for (q0, qr, b0, br, o0) in split]   # partials live behind row nq
            rem = len(split) // chunks if split else 0
            rows = nq + chunks * rem * _PAIR_ROWS if split else nq
            plan = (key, _native.knn_items_from_host(direct + split, dev), len(direct), len(split), chunks, chunk_rows,
                    torch.empty((rows, self.k), dtype=torch.float32, device=dev),
                    torch.empty((rows, self.k), dtype=torch.int64, device=dev),
                    torch.zeros(2, dtype=torch.int64, device=dev))    # sweep barrier: arrivals, give-up flag
            self._pair_plan = plan
        _, items, n_direct, n_split, chunks, chunk_rows, buf_d, buf_i, sync = plan
        if n_split == 0:
            _native.knn_search_pair_items(q16, qn, self._bank16, self._bank_sqnorm, items, n_direct, metric, self.k,
                                          dist, idx, sync, _SWEEP_SYNC_TILES)
            return dist, idx
        _native.knn_search_pair_items(q16, qn, self._bank16, self._bank_sqnorm, items, n_direct + n_split, metric,
                                      self.k, buf_d, buf_i, sync, _SWEEP_SYNC_TILES)
        q_pad = (n_split // chunks) * _PAIR_ROWS
        first_tail = n_direct * _PAIR_ROWS
        _native.knn_merge(buf_d[nq:], buf_i[nq:], chunks, q_pad, nq - first_tail, self.k, chunk_rows, nb, metric == 1,
                          buf_d[first_tail:], buf_i[first_tail:])
        return buf_d[:nq], buf_i[:nq]

    def serialize_index(self) -> None:
        pass

    def deserialize_index(self) -> None:
        pass
