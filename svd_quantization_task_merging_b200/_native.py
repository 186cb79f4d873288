"""ctypes binding of libsvdq.so (C ABI in include/svdq.h).

There is no CPU fallback: if the library is missing or no CUDA device is visible, every
compute entry point raises.  Return codes follow include/svdq.h: negative -> ValueError
(the reference raises ValueError for the same misuse), positive -> RuntimeError (CUDA error).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsvdq.so")
ABI_VERSION = 7

DTYPE_CODE = {"float32": 0, "bfloat16": 1, "float16": 2}
STRATEGY_CODE = {"union": 0, "intersection": 1, "majority": 2}

_lib: Optional[C.CDLL] = None

_vp, _i32, _i64, _f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float

_SIGNATURES = {
    "svdq_abi_version": (C.c_int, []),
    "svdq_last_error": (C.c_char_p, []),
    "svdq_k4_scratch_bytes": (C.c_int64, []),
    "svdq_tv_mask_gram": (C.c_int, [_i32, _i32, _i32, _i32, _i64, _i32] + [_vp] * 10),
    "svdq_tv_mask_gram_bits": (C.c_int, [_i32, _i32, _i32, _i32, _i64, _i32] + [_vp] * 10),
    "svdq_host_pack_mask": (C.c_int, [_vp, _i64, _vp, _i32]),
    "svdq_host_pack_mask_batch": (C.c_int, [_vp, _vp, _vp, _i64, _i32]),
    "svdq_host_kmeans": (C.c_int, [_vp, _i32, _i32, _i32, C.c_uint32, _i32, _i32, C.c_double, _vp, _vp]),
    "svdq_mask_pack": (C.c_int, [_i32, _i32, _i64, _i32] + [_vp] * 8),
    "svdq_gram_staged": (C.c_int, [_i32, _i32, _i32, _i64, _i32] + [_vp] * 9),
    "svdq_gram_reduce": (C.c_int, [_i32, _i32, _i64, _i32] + [_vp] * 11),
    "svdq_param_solve": (C.c_int, [_i32, _i64, _i32, _f32, _i32, _i32, _i32, _i32] + [_vp] * 22),
    "svdq_param_average": (C.c_int, [_i32, _i64] + [_vp] * 12),
    "svdq_project_exact": (C.c_int, [_i32, _i32, _i32, _i32, _i32, _i64, _i32] + [_vp] * 11),
    "svdq_param_requantize": (C.c_int, [_i32, _i64, _i32, _i32] + [_vp] * 12),
    "svdq_reconstruct_merge": (C.c_int, [_i32, _i32, _i32, _i32, _i32, _i64, _i32] + [_vp] * 20 + [_f32, _vp]),
    "svdq_reconstruct_merge_basis": (C.c_int, [_i32, _i32, _i32, _i32, _i64, _i32] + [_vp] * 20),
    "svdq_diag_finalize": (C.c_int, [_i32, _i64] + [_vp] * 7),
    "svdq_mask_tile_counts": (C.c_int, [_i64, _i32] + [_vp] * 8),
    "svdq_reload_merge": (C.c_int, [_i32, _i32, _i32, _f32, _i64, _i32] + [_vp] * 14),
    "svdq_basis_offsets": (C.c_int, [_i64, _i32, _i32] + [_vp] * 5),
    "svdq_write_basis": (C.c_int, [_i32, _i32, _i32, _i32, _i32, _i64, _i32] + [_vp] * 14),
    "svdq_rtvq_quantize": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _i64, _i32] + [_vp] * 5 + [_i64, _vp]),
    "svdq_rtvq_dequantize": (C.c_int, [_vp, _i64, _i32, _i32, _i64] + [_vp] * 4),
    "svdq_absmax_quantize": (C.c_int, [_vp, _i64, _i32, _vp, _i32] + [_vp] * 3),
    "svdq_combine_masks": (C.c_int, [_vp, _i32, _i64, _i32, _vp, _vp]),
    "svdq_unpack_mask": (C.c_int, [_vp, _i64, _vp, _vp]),
    "svdq_project_scratch_bytes": (C.c_size_t, []),
    "svdq_select_chunk_elems": (C.c_int, []),
    "svdq_basis_project": (C.c_int, [_i32, _vp, _i64, _i32, _i64] + [_vp] * 5),
    "svdq_basis_expand": (C.c_int, [_i32, _vp, _i64, _i32, _vp, _i64, _i32, _i64, _vp, _vp, _vp, _f32, _vp, _vp]),
    "svdq_mask_offsets": (C.c_int, [_vp, _i64, _i32, _vp, _vp]),
    "svdq_mask_select": (C.c_int, [_vp, _i32, _vp, _i64, _i32, _vp, _vp, _vp]),
    "svdq_mask_scatter": (C.c_int, [_vp, _i32, _vp, _i64, _i32, _vp, _vp, _vp]),
}

EXPORTS = tuple(_SIGNATURES.keys())


class NativeLibraryError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load libsvdq.so (built in-tree by build.py).  Raises if it is missing: no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            f"{LIB_PATH} not found: build it with `python -m svd_quantization_task_merging_b200.build` "
            "(nvcc, sm_100a).  There is no CPU fallback for the SVD-Hybrid merge path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.svdq_abi_version() != ABI_VERSION:
        raise NativeLibraryError(f"libsvdq.so ABI {lib.svdq_abi_version()} != expected {ABI_VERSION}")
    _lib = lib
    return lib


def require_cuda():
    import torch
    load()
    if not torch.cuda.is_available():
        raise NativeLibraryError("no CUDA device visible: the SVD-Hybrid merge path runs only on the GPU "
                                 "(sm_100a kernels in libsvdq.so); there is no CPU fallback")


def check(rc: int, what: str = ""):
    if rc == 0:
        return
    msg = load().svdq_last_error().decode("utf-8", "replace")
    if rc < 0:
        raise ValueError(msg or what)
    raise RuntimeError(msg or f"{what}: CUDA error {rc}")


def call(name: str, *args):
    check(getattr(load(), name)(*args), name)


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
