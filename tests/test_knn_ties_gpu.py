"""Exact tie / ordering behaviour of the k-NN top-k epilogue (GPU, through the C ABI).

Integer-valued fp16 vectors make every distance an exactly representable integer, so the CUDA kernels must reproduce
the reference order bit for bit: ascending distance, ties to the LOWER index (faiss's heap with `CMax<float, int64_t>`
keeps the first of equal candidates; utils/knn_util.py:65-106 is what is asked of it), across the 16-column candidate
groups, the 32-column chunks, the two column halves of a tile, the tiles of an item, the bank slices of the split path
and the cta_group::2 pair kernel.  The inputs are adversarial for the candidate lists of `insert_candidates`
(csrc/knn_tcgen05.cu): thousands of exact ties, and a bank whose distances DEScend with the index so that every
column is a candidate for every row.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _exact_topk(q: torch.Tensor, bank: torch.Tensor, k: int):
    """Integer arithmetic in int64, stable sort: ascending distance, ties to the lower index."""
    qi, bi = q.to(torch.int64), bank.to(torch.int64)
    d = (qi * qi).sum(1, keepdim=True) + (bi * bi).sum(1)[None, :] - 2 * qi @ bi.T
    order = torch.sort(d, dim=1, stable=True)
    return order.values[:, :k].float(), order.indices[:, :k]


def _search(q, bank, k):
    from foundpose_b200.utils import knn_util

    index = knn_util.KNN(k=k, metric="l2")
    index.fit(bank.cuda())
    d, i = index.search(q.cuda())
    return d.cpu(), i.cpu()


@pytest.mark.parametrize("nq,nb,dim,k", [(256, 5000, 64, 5), (200, 3000, 64, 16), (130, 777, 128, 3), (64, 4097, 64, 8),
                                         (100, 40000, 64, 5),       # bank split over all SMs + merge
                                         (9500, 9000, 64, 5),       # pair kernel
                                         (19000, 8448, 64, 16)])    # pair kernel: whole wave + split tail wave, k = 16
def test_integer_distances_with_many_ties_are_ordered_like_the_reference(nq, nb, dim, k):
    g = torch.Generator().manual_seed(nq * 7 + nb)
    bank = torch.randint(-2, 3, (nb, dim), generator=g).float()
    q = torch.randint(-2, 3, (nq, dim), generator=g).float()
    d, i = _search(q, bank, k)
    rd, ri = _exact_topk(q, bank, k)
    assert torch.equal(i, ri), "ids differ from the stable exact order (ties must go to the lower index)"
    assert torch.equal(d, rd), "integer distances must be exact"


@pytest.mark.parametrize("k", [2, 3, 5, 8, 16])
def test_descending_distances_make_every_column_a_candidate(k):
    # row j = (nb - j, 0, ...): the distance to the zero query falls with j, so each bank row displaces the current best
    nb, dim, nq = 2048, 64, 130
    bank = torch.zeros(nb, dim)
    bank[:, 0] = torch.arange(nb, 0, -1, dtype=torch.float32)
    q = torch.zeros(nq, dim)
    q[:, 1] = torch.arange(nq, dtype=torch.float32) % 7          # row-dependent constant offset, order unchanged
    d, i = _search(q, bank, k)
    rd, ri = _exact_topk(q, bank, k)
    assert torch.equal(i, ri) and torch.equal(d, rd)
    assert torch.equal(i[0], torch.arange(nb - 1, nb - 1 - k, -1))


@pytest.mark.parametrize("k", [3, 5, 16])
def test_identical_bank_rows_return_the_first_k_indices(k):
    nb, dim, nq = 3000, 64, 129
    bank = torch.ones(nb, dim)
    q = torch.zeros(nq, dim)
    d, i = _search(q, bank, k)
    assert torch.equal(i, torch.arange(k).expand(nq, k))
    assert torch.equal(d, torch.full((nq, k), float(dim)))
