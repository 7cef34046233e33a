"""CPU oracle of the FoundPose per-crop hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in plain fp32 PyTorch-on-CPU / numpy, the arithmetic the reference
(facebookresearch/foundpose @ 3103473b, external/dinov2 @ e1277af) performs on the path

    crop -> DINOv2 ViT patch features -> grid sampling -> PCA projection -> k-NN vs the object's
    template bank -> tf-idf bag-of-words template retrieval -> cyclic 2D-3D correspondences.

Each function cites the reference file:line it follows.  Only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s cpu_baseline / `--impl reference` legs may import it - as the checker or as the
timed CPU baseline, never as part of the shipped product path (`foundpose_b200/` does not import
`oracle`; the product raises ImportError when libfoundpose_b200.so is missing).

How the oracle is pinned
------------------------
The reference ships no tests, golden vectors or fixtures for this path, and the
k-NN arithmetic lives in faiss 1.8.0 (conda_foundpose_gpu.yaml:19), which is neither vendored in
the reference nor installed in this image.  The oracle is therefore pinned against OUTPUTS OF THE
REFERENCE ITSELF: `tests/golden/make_golden.py` (run in the build container, where
/root/reference exists) imports the reference's own modules (utils/dinov2_utils.py,
feature_util.py, projector_util.py, knn_util.py, template_util.py, corresp_util.py and
external/dinov2) on seeded synthetic inputs and stores their outputs under `tests/golden/`;
`tests/test_oracle_golden.py` checks every oracle function against those files.  For faiss the
generator injects a stub whose IndexFlatL2 / IndexFlatIP follow faiss 1.8.0's published
algorithm (exhaustive_L2sqr_blas: ||x||^2 + ||y||^2 - 2<x,y> clamped at 0; exhaustive inner
product), so the k-NN leg is pinned only to that restatement - "parity unpinned against faiss
binaries" is stated here and in DESIGN.md.

Canonical rules added where the reference is under-determined
-----------------------------------------------------------------------------
* k-NN / top-k ties are broken by ascending index.
* index equality is asserted only where the oracle's own margin to the next candidate exceeds
  1e-4 relative (`knn.topk_margin`).
"""

from . import corresp, feature, knn, pca, template, vit  # noqa: F401
