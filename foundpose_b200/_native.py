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
