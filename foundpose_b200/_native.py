"""ctypes binding of libfoundpose_b200.so (the C ABI declared in include/foundpose_b200.h).

The library is built in-tree by `__graft_entry__.build()` (or `make -C foundpose_b200/csrc`).
There is NO CPU or PyTorch fallback behind this module: if the shared library is missing or a
symbol cannot be resolved the import of the product path fails loudly.
"""

from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Optional

import torch

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libfoundpose_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_PKG_DIR), "include", "foundpose_b200.h")

_lib: Optional[ctypes.CDLL] = None


class NativeError(RuntimeError):
    """Raised when an fp_* entry point returns a non-zero status."""


def declared_symbols(header_path: str = HEADER_PATH) -> List[str]:
    """Names of every function declared in include/foundpose_b200.h."""
    with open(header_path, "r", encoding="utf-8") as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fp_[a-z0-9_]+)\s*\(", text)))


def load() -> ctypes.CDLL:
    """Loads the shared library (once) and checks that every declared symbol is exported."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing. Build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `make -C foundpose_b200/csrc`. foundpose_b200 has no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    if missing:
        raise ImportError(f"{LIB_PATH} does not export: {missing}")
    lib.fp_last_error.restype = ctypes.c_char_p
    lib.fp_last_error.argtypes = []
    lib.fp_version.restype = ctypes.c_int
    lib.fp_launch_count.restype = ctypes.c_ulonglong
    lib.fp_launch_count_category.restype = ctypes.c_ulonglong
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().fp_last_error().decode("utf-8", "replace")
        raise NativeError(f"{what} failed with status {status}: {msg}")


def ptr(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    """Device pointer of a tensor (NULL for None)."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr(device: Optional[torch.device] = None) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t: torch.Tensor, name: str, dtype: Optional[torch.dtype] = None) -> None:
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (foundpose_b200 has no CPU path)")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise ValueError(f"{name} must have dtype {dtype}, got {t.dtype}")


# ---------------------------------------------------------------------------------------------
# Thin typed wrappers (one per C entry point). Higher-level mirrors of the reference API live in
# foundpose_b200/utils/.
# ---------------------------------------------------------------------------------------------
EPI_BIAS_F16 = 0
EPI_BIAS_GELU_F16 = 1
EPI_RESID_F32 = 2
EPI_BIAS_F32 = 4


def gemm_tn_f16(
    a: torch.Tensor,
    b: torch.Tensor,
    epilogue: int,
    bias: Optional[torch.Tensor] = None,
    gamma: Optional[torch.Tensor] = None,
    out_f16: Optional[torch.Tensor] = None,
    out_f32: Optional[torch.Tensor] = None,
) -> None:
    """C = A[M,K] . B[N,K]^T with a fused epilogue (see include/foundpose_b200.h)."""
    lib = load()
    require_cuda(a, "a", torch.float16)
    require_cuda(b, "b", torch.float16)
    m, k = a.shape
    n = b.shape[0]
    assert b.shape[1] == k
    check(
        lib.fp_gemm_tn_f16(
            ctypes.c_int(epilogue), ptr(a), ctypes.c_int(a.stride(0)), ptr(b),
            ctypes.c_int(b.stride(0)), ctypes.c_int(m), ctypes.c_int(n), ctypes.c_int(k),
            ptr(bias), ptr(gamma), ptr(out_f16),
            ctypes.c_int(out_f16.stride(0) if out_f16 is not None else 0), ptr(out_f32),
            ctypes.c_int(out_f32.stride(0) if out_f32 is not None else 0), stream_ptr(a.device),
        ),
        "fp_gemm_tn_f16",
    )


def umma_probe(a: torch.Tensor, b: torch.Tensor, b_mn_major: bool) -> torch.Tensor:
    lib = load()
    out = torch.empty(128, 64, dtype=torch.float32, device=a.device)
    check(
        lib.fp_umma_probe(ptr(a), ptr(b), ptr(out), ctypes.c_int(int(b_mn_major)),
                          stream_ptr(a.device)),
        "fp_umma_probe",
    )
    return out


def _i(v) -> ctypes.c_int:
    return ctypes.c_int(int(v))


def _l(v) -> ctypes.c_int64:
    return ctypes.c_int64(int(v))


def _f(v) -> ctypes.c_float:
    return ctypes.c_float(float(v))


def call(name: str, *args) -> None:
    """Calls lib.<name>(*args) and raises NativeError on a non-zero status."""
    check(getattr(load(), name)(*args), name)


KNN_ITEM_BYTES = 32  # sizeof(fp_knn_item)


def layernorm_f16(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    require_cuda(x, "x", torch.float32)
    m, d = x.shape
    y = torch.empty((m, d), dtype=torch.float16, device=x.device)
    call("fp_layernorm_f16", ptr(x), ptr(y), ptr(weight), ptr(bias), _i(m), _i(d), _f(eps), stream_ptr(x.device))
    return y


def attention_f16(qkv: torch.Tensor, batch: int, tokens: int, heads: int) -> torch.Tensor:
    require_cuda(qkv, "qkv", torch.float16)
    assert qkv.shape == (batch * tokens, 3 * heads * 64)
    out = torch.empty((batch * tokens, heads * 64), dtype=torch.float16, device=qkv.device)
    call("fp_attention_f16", ptr(qkv), ptr(out), _i(batch), _i(tokens), _i(heads), stream_ptr(qkv.device))
    return out


def convert_rows_f16(x: torch.Tensor, l2_normalize: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    require_cuda(x, "x", torch.float32)
    rows, dim = x.shape
    if out is None:
        out = torch.empty((rows, dim), dtype=torch.float16, device=x.device)
    call("fp_convert_rows_f16", ptr(x), ptr(out), _l(rows), _i(dim), _i(int(l2_normalize)), stream_ptr(x.device))
    return out


def unit_rows_f16(x16: torch.Tensor, out: torch.Tensor, sqnorm: torch.Tensor) -> None:
    """f16 rows -> unit f16 rows + their squared norms (fp_unit_rows_f16)."""
    require_cuda(x16, "x16", torch.float16)
    call("fp_unit_rows_f16", ptr(x16), ptr(out), ptr(sqnorm), _l(x16.shape[0]), _i(x16.shape[1]), stream_ptr(x16.device))


def split_rows_f16(x: torch.Tensor, pattern: int, l2_normalize: bool, scale: float,
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 [rows, dim] -> f16 [rows, 3*dim] hi/lo split (fp_split_rows_f16)."""
    require_cuda(x, "x", torch.float32)
    rows, dim = x.shape
    if out is None:
        out = torch.empty((rows, 3 * dim), dtype=torch.float16, device=x.device)
    call("fp_split_rows_f16", ptr(x), ptr(out), _l(rows), _i(dim), _i(pattern), _i(int(l2_normalize)), _f(scale),
         stream_ptr(x.device))
    return out


def row_sqnorm_f16(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    require_cuda(x, "x", torch.float16)
    rows, dim = x.shape
    if out is None:
        out = torch.empty((rows,), dtype=torch.float32, device=x.device)
    call("fp_row_sqnorm_f16", ptr(x), ptr(out), _l(rows), _i(dim), stream_ptr(x.device))
    return out


def knn_num_items(q_rows: int) -> int:
    return int(load().fp_knn_num_items(_i(q_rows)))


def new_knn_items(num_items: int, device) -> torch.Tensor:
    """Device buffer holding `num_items` fp_knn_item structs (as int64 words)."""
    return torch.zeros((max(num_items, 1), KNN_ITEM_BYTES // 8), dtype=torch.int64, device=device)


def knn_items_dense(items: torch.Tensor, q_total: int, b_row0: int, b_rows: int) -> None:
    call("fp_knn_items_dense", ptr(items), _i(q_total), _i(b_row0), _i(b_rows), stream_ptr(items.device))


def knn_items_split(items: torch.Tensor, q_total: int, b_row0: int, b_rows: int, num_chunks: int, chunk_rows: int) -> None:
    call("fp_knn_items_split", ptr(items), _i(q_total), _i(b_row0), _i(b_rows), _i(num_chunks), _i(chunk_rows),
         stream_ptr(items.device))


def knn_merge(part_d, part_i, num_chunks: int, q_pad: int, nq: int, k: int, chunk_rows: int, b_rows: int,
              descending: bool, out_d, out_i) -> None:
    call("fp_knn_merge", ptr(part_d), ptr(part_i), _i(num_chunks), _i(q_pad), _i(nq), _i(k), _i(chunk_rows),
         _i(b_rows), _i(int(descending)), ptr(out_d), ptr(out_i), stream_ptr(part_d.device))


def knn_search_items(q16, q_sqnorm, bank16, bank_sqnorm, items, num_items, metric: int, k: int, out_d, out_i) -> None:
    require_cuda(q16, "q16", torch.float16)
    require_cuda(bank16, "bank16", torch.float16)
    assert q16.shape[1] == bank16.shape[1]
    call("fp_knn_search_items", ptr(q16), _l(q16.shape[0]), ptr(q_sqnorm), ptr(bank16), _l(bank16.shape[0]),
         ptr(bank_sqnorm), _i(q16.shape[1]), ptr(items), _i(num_items), _i(metric), _i(k), ptr(out_d), ptr(out_i),
         stream_ptr(q16.device))


def knn_search_pair_items(q16, q_sqnorm, bank16, bank_sqnorm, items, num_items, metric: int, k: int, out_d, out_i,
                          sync_counter: Optional[torch.Tensor] = None, sync_tiles: int = 0) -> None:
    """Pair kernel (cta_group::2, queries resident in shared memory); items hold up to 256 query rows.
    sync_counter (int64 [2], device) + sync_tiles > 0 enable the sweep barrier of items whose 4th word is set."""
    require_cuda(q16, "q16", torch.float16)
    require_cuda(bank16, "bank16", torch.float16)
    assert q16.shape[1] == bank16.shape[1]
    call("fp_knn_search_pair_items", ptr(q16), _l(q16.shape[0]), ptr(q_sqnorm), ptr(bank16), _l(bank16.shape[0]),
         ptr(bank_sqnorm), _i(q16.shape[1]), ptr(items), _i(num_items), _i(metric), _i(k), ptr(out_d), ptr(out_i),
         ptr(sync_counter), _i(sync_tiles if sync_counter is not None else 0), stream_ptr(q16.device))


def knn_items_from_host(rows, device) -> torch.Tensor:
    """Device fp_knn_item array from host tuples (q_row0, q_rows, b_row0, b_rows, out_row0[, sweep participants])."""
    import numpy as np

    arr = np.zeros((max(len(rows), 1), 4), dtype=np.int64)
    for n, row in enumerate(rows):
        q0, qr, b0, br, o0 = row[:5]
        arr[n, 0] = (int(q0) & 0xFFFFFFFF) | (int(qr) << 32)
        arr[n, 1] = (int(b0) & 0xFFFFFFFF) | (int(br) << 32)
        arr[n, 2] = int(o0)
        arr[n, 3] = int(row[5]) if len(row) > 5 else 0
    return torch.from_numpy(arr).to(device)


def num_sms() -> int:
    return int(load().fp_num_sms())


def pca_project(x16, comp16, bias, out_f32, out_f16=None) -> None:
    require_cuda(x16, "x16", torch.float16)
    require_cuda(comp16, "comp16", torch.float16)
    m, dd = x16.shape
    d = comp16.shape[0]
    call("fp_pca_project", ptr(x16), ptr(comp16), ptr(bias), _i(m), _i(dd), _i(d), ptr(out_f32), ptr(out_f16),
         stream_ptr(x16.device))


CROP_PARAM_STRIDE = 40


def crop_warp(images: Optional[torch.Tensor], masks_u8: Optional[torch.Tensor], params: torch.Tensor,
              crop_w: int, crop_h: int, out_images: Optional[torch.Tensor] = None,
              out_masks: Optional[torch.Tensor] = None, out_boxes: Optional[torch.Tensor] = None):
    """images [n_img, H, W, C] uint8 or fp32, masks [B, H, W] uint8, params [B, 40] fp64 (see the header)."""
    require_cuda(params, "params", torch.float64)
    b = params.shape[0]
    assert params.shape[1] == CROP_PARAM_STRIDE
    dev = params.device
    n_img = c = 0
    is_f32 = 0
    if images is not None:
        assert images.dtype in (torch.uint8, torch.float32) and images.is_cuda and images.is_contiguous()
        n_img, h, w, c = images.shape
        is_f32 = int(images.dtype == torch.float32)
        if out_images is None:
            out_images = torch.empty((b, c, crop_h, crop_w), dtype=torch.float32, device=dev)
    box_ws = None
    if masks_u8 is not None:
        require_cuda(masks_u8, "masks", torch.uint8)
        assert masks_u8.shape[0] == b
        h, w = masks_u8.shape[1:]
        if out_masks is None:
            out_masks = torch.empty((b, crop_h, crop_w), dtype=torch.uint8, device=dev)
        if out_boxes is None:
            out_boxes = torch.empty((b, 4), dtype=torch.float32, device=dev)
        box_ws = torch.empty((4 * b,), dtype=torch.int32, device=dev)
    if images is not None and masks_u8 is not None:
        assert tuple(images.shape[1:3]) == tuple(masks_u8.shape[1:]), "image / mask size mismatch"
    call("fp_crop_warp", ptr(images), _i(is_f32), _i(n_img), _i(h), _i(w), _i(c), ptr(masks_u8), ptr(params), _i(b),
         _i(crop_w), _i(crop_h), ptr(out_images), ptr(out_masks), ptr(out_boxes), ptr(box_ws), stream_ptr(dev))
    return out_images, out_masks, out_boxes


def filter_points_by_mask(points, masks_u8, out_points, out_ids, out_counts) -> None:
    require_cuda(points, "points", torch.float32)
    require_cuda(masks_u8, "masks", torch.uint8)
    b, h, w = masks_u8.shape
    call("fp_filter_points_by_mask", ptr(points), _i(points.shape[0]), ptr(masks_u8), _i(b), _i(h), _i(w),
         ptr(out_points), ptr(out_ids), ptr(out_counts), _i(out_points.shape[1]), stream_ptr(points.device))


def sample_features(tokens, hp, wp, points, counts, img_w, img_h, out_f32=None, out_f16=None) -> None:
    require_cuda(tokens, "tokens", torch.float32)
    b = tokens.shape[0]
    c = tokens.shape[-1]
    stride = points.shape[1]
    call("fp_sample_features", ptr(tokens), _i(b), _i(hp), _i(wp), _i(c), ptr(points), ptr(counts), _i(stride),
         _f(img_w), _f(img_h), ptr(out_f32), ptr(out_f16), stream_ptr(tokens.device))


def calc_tfidf(word_ids, word_dists, row_start, row_count, idf, soft: bool, sigma2: float, sqrt_input: bool, out) -> None:
    require_cuda(word_ids, "word_ids", torch.int64)
    require_cuda(word_dists, "word_dists", torch.float32)
    call("fp_calc_tfidf", ptr(word_ids), ptr(word_dists), _i(word_ids.shape[1]), ptr(row_start), ptr(row_count),
         _i(out.shape[0]), ptr(idf), _i(out.shape[1]), _i(int(soft)), _f(sigma2), _i(int(sqrt_input)), ptr(out),
         stream_ptr(out.device))


def row_norm_f32(x, out=None) -> torch.Tensor:
    require_cuda(x, "x", torch.float32)
    if out is None:
        out = torch.empty((x.shape[0],), dtype=torch.float32, device=x.device)
    call("fp_row_norm_f32", ptr(x), ptr(out), _i(x.shape[0]), _i(x.shape[1]), stream_ptr(x.device))
    return out


def bow_scores(descs, desc_norm, q, out) -> None:
    require_cuda(descs, "descs", torch.float32)
    require_cuda(q, "q", torch.float32)
    call("fp_bow_scores", ptr(descs), ptr(desc_norm), ptr(q), _i(descs.shape[0]), _i(q.shape[0]), _i(descs.shape[1]),
         ptr(out), stream_ptr(q.device))


def topk_rows(x, k: int, out_v, out_i) -> None:
    require_cuda(x, "x", torch.float32)
    call("fp_topk_rows", ptr(x), _i(x.shape[0]), _i(x.shape[1]), _i(k), ptr(out_v), ptr(out_i), stream_ptr(x.device))


def build_pair_items(top_ids, topn, tpl_off, q_start, q_count, max_q, max_p, items_q2o, items_o2q) -> None:
    call("fp_build_pair_items", ptr(top_ids), _i(top_ids.numel()), _i(topn), ptr(tpl_off), ptr(q_start), ptr(q_count),
         _i(max_q), _i(max_p), ptr(items_q2o), ptr(items_o2q), stream_ptr(top_ids.device))


def kmeans_update(samples: torch.Tensor, assign: torch.Tensor, k: int, sums: torch.Tensor, counts: torch.Tensor,
                  centroids: torch.Tensor) -> None:
    require_cuda(samples, "samples", torch.float32)
    require_cuda(assign, "assign", torch.int64)
    require_cuda(centroids, "centroids", torch.float32)
    n, d = samples.shape
    assert sums.dtype == torch.int64 and sums.numel() == k * d and counts.dtype == torch.int32 and counts.numel() == k
    call("fp_kmeans_update", ptr(samples), ptr(assign), _l(n), _i(d), _i(k), ptr(sums), ptr(counts), ptr(centroids),
         stream_ptr(samples.device))


def pnp_ransac(coord_2d, coord_3d, counts, intrinsics, iters: int, thresh: float, confidence: float, seed: int,
               problem_offset: int = 0):
    """fp_pnp_ransac over P problems; returns a dict of freshly allocated device tensors."""
    require_cuda(coord_2d, "coord_2d", torch.float32)
    require_cuda(coord_3d, "coord_3d", torch.float32)
    require_cuda(counts, "counts", torch.int32)
    require_cuda(intrinsics, "intrinsics", torch.float64)
    P, M = coord_2d.shape[0], coord_2d.shape[1]
    if coord_3d.shape != (P, M, 3) or coord_2d.shape != (P, M, 2) or counts.shape != (P,) or intrinsics.shape != (P, 4):
        raise ValueError("pnp_ransac: expected coord_2d [P,M,2], coord_3d [P,M,3], counts [P], intrinsics [P,4]")
    dev = coord_2d.device
    out = {
        "success": torch.zeros(P, dtype=torch.int32, device=dev),
        "R": torch.zeros(P, 3, 3, dtype=torch.float64, device=dev),
        "t": torch.zeros(P, 3, dtype=torch.float64, device=dev),
        "inlier_mask": torch.zeros(P, M, dtype=torch.uint8, device=dev),
        "num_inliers": torch.zeros(P, dtype=torch.int32, device=dev),
        "iters_run": torch.zeros(P, dtype=torch.int32, device=dev),
        "best_hyp": torch.zeros(P, dtype=torch.int32, device=dev),
    }
    call("fp_pnp_ransac", ptr(coord_2d), ptr(coord_3d), ptr(counts), ptr(intrinsics), _i(P), _i(M), _i(iters),
         ctypes.c_double(thresh), ctypes.c_double(confidence), ctypes.c_uint64(seed & 0xFFFFFFFFFFFFFFFF),
         _i(problem_offset), ptr(out["success"]), ptr(out["R"]), ptr(out["t"]), ptr(out["inlier_mask"]),
         ptr(out["num_inliers"]), ptr(out["iters_run"]), ptr(out["best_hyp"]), stream_ptr(dev))
    return out


def cyclic_buddies_workspace(num_pairs: int, max_q: int, top_k: int, device) -> Optional[torch.Tensor]:
    lib = load()
    lib.fp_cyclic_buddies_workspace_bytes.restype = ctypes.c_uint64
    nbytes = int(lib.fp_cyclic_buddies_workspace_bytes(_i(num_pairs), _i(max_q), _i(top_k)))
    if nbytes == 0:
        return None
    return torch.empty(((nbytes + 7) // 8,), dtype=torch.int64, device=device)


def cyclic_buddies(points, q_start, q_count, q2o, o2q, top_ids, topn, tpl_off, feat_perm, vertices, max_q, max_p,
                   top_k, out_qids, out_vids, out_dists, out_scores, out_c2d, out_c3d, out_count,
                   workspace: Optional[torch.Tensor] = None) -> None:
    if workspace is None:
        workspace = cyclic_buddies_workspace(top_ids.numel(), max_q, top_k, points.device)
    ws_bytes = workspace.numel() * 8 if workspace is not None else 0
    call("fp_cyclic_buddies", ptr(points), ptr(q_start), ptr(q_count), ptr(q2o), ptr(o2q), ptr(top_ids),
         _i(top_ids.numel()), _i(topn), ptr(tpl_off), ptr(feat_perm), ptr(vertices), _i(max_q), _i(max_p), _i(top_k),
         ptr(out_qids), ptr(out_vids), ptr(out_dists), ptr(out_scores), ptr(out_c2d), ptr(out_c3d), ptr(out_count),
         ptr(workspace), ctypes.c_uint64(ws_bytes), stream_ptr(points.device))


# ---- ViT handle ---------------------------------------------------------------------------------
class VitConfig(ctypes.Structure):
    _fields_ = [("embed_dim", ctypes.c_int), ("num_heads", ctypes.c_int), ("num_blocks", ctypes.c_int),
                ("num_register_tokens", ctypes.c_int), ("patch_size", ctypes.c_int), ("img_h", ctypes.c_int),
                ("img_w", ctypes.c_int), ("fuse_layernorm", ctypes.c_int)]


class VitWeights(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in
                ("patch_w", "patch_b", "cls_pos", "reg_tokens", "pos_patch", "norm_w", "norm_b")]


class VitBlockWeights(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in
                ("norm1_w", "norm1_b", "qkv_w", "qkv_b", "proj_w", "proj_b", "ls1", "norm2_w", "norm2_b",
                 "fc1_w", "fc1_b", "fc2_w", "fc2_b", "ls2", "qkv_colsum", "fc1_colsum")]


def vit_create(cfg: VitConfig, weights: VitWeights, blocks, max_batch: int) -> ctypes.c_void_p:
    lib = load()
    arr = (VitBlockWeights * len(blocks))(*blocks)
    handle = ctypes.c_void_p(0)
    check(lib.fp_vit_create(ctypes.byref(cfg), ctypes.byref(weights), arr, _i(max_batch), ctypes.byref(handle)),
          "fp_vit_create")
    return handle


def vit_destroy(handle: ctypes.c_void_p) -> None:
    lib = load()
    lib.fp_vit_destroy.restype = None
    lib.fp_vit_destroy(handle)


def vit_forward(handle, images, layer: int, facet: int, apply_norm: bool, out_tokens, out_tokens_f16, out_cls) -> None:
    require_cuda(images, "images", torch.float32)
    call("fp_vit_forward", handle, ptr(images), _i(images.shape[0]), _i(layer), _i(facet), _i(int(apply_norm)),
         ptr(out_tokens), ptr(out_tokens_f16), ptr(out_cls), stream_ptr(images.device))
