"""ctypes binding for the plain-C oracle (oracle/rtvq_ref.c).  TEST INFRASTRUCTURE."""
import ctypes as C
from typing import List, Optional, Tuple

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build())
        _lib.ref_combine_masks.restype = C.c_int
        _lib.ref_asym_quant.restype = C.c_int
        _lib.ref_rtvq.restype = C.c_int
        _lib.ref_select_rank.restype = C.c_int
        _lib.ref_absmax_quant.restype = C.c_int
        _lib.ref_asym_dequant.restype = None
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


STRATEGY = {"union": 0, "intersection": 1, "majority": 2}


def combine_masks(masks: List[np.ndarray], strategy: str) -> np.ndarray:
    if strategy not in STRATEGY:
        raise ValueError(f"Unknown mask strategy: {strategy}")
    if not masks:
        raise ValueError("Empty mask list")
    flat = [np.ascontiguousarray(m, dtype=np.uint8).reshape(-1) for m in masks]
    n = flat[0].size
    ptrs = (C.POINTER(C.c_uint8) * len(flat))(*[m.ctypes.data_as(C.POINTER(C.c_uint8)) for m in flat])
    out = np.empty(n, np.uint8)
    rc = lib().ref_combine_masks(ptrs, C.c_int(len(flat)), C.c_int64(n), C.c_int(STRATEGY[strategy]),
                                 out.ctypes.data_as(C.POINTER(C.c_uint8)))
    assert rc == 0
    return out.astype(bool).reshape(np.shape(masks[0]))


def asym_quant(x: np.ndarray, bits: int) -> Tuple[np.ndarray, np.float32, np.float32]:
    x = np.ascontiguousarray(x, np.float32).reshape(-1)
    q = np.empty(x.size, np.int32)
    s, z = C.c_float(), C.c_float()
    rc = lib().ref_asym_quant(_fp(x), C.c_int64(x.size), C.c_int(bits), _ip(q), C.byref(s), C.byref(z))
    assert rc == 0
    return q, np.float32(s.value), np.float32(z.value)


def asym_dequant(q: np.ndarray, scale, zp) -> np.ndarray:
    q = np.ascontiguousarray(q, np.int32).reshape(-1)
    out = np.empty(q.size, np.float32)
    lib().ref_asym_dequant(_ip(q), C.c_int64(q.size), C.c_float(float(scale)), C.c_float(float(zp)), _fp(out))
    return out


def rtvq(x: np.ndarray, bits: int, stages: int):
    """-> (codes[int32 stages x n], scale[stages], zp[stages], resnorm[stages], deq_sum[n])"""
    x = np.ascontiguousarray(x, np.float32).reshape(-1)
    n = x.size
    codes = np.zeros((stages, n), np.int32)
    scale = np.zeros(stages, np.float32)
    zp = np.zeros(stages, np.float32)
    rn = np.zeros(stages, np.float32)
    deq = np.zeros(n, np.float32)
    rc = lib().ref_rtvq(_fp(x), C.c_int64(n), C.c_int(bits), C.c_int(stages), _ip(codes), _fp(scale), _fp(zp),
                        _fp(rn), _fp(deq))
    assert rc >= 0
    return codes, scale, zp, rn, deq


def select_rank(S: np.ndarray, thr: float, max_rank: Optional[int] = None, min_rank: int = 1):
    S = np.ascontiguousarray(S, np.float32).reshape(-1)
    cum = np.zeros(S.size, np.float32)
    k = lib().ref_select_rank(_fp(S), C.c_int(S.size), C.c_float(thr), C.c_int(max_rank or 0), C.c_int(min_rank),
                              _fp(cum))
    return int(k), cum


def absmax_quant(x: np.ndarray, bits: int):
    x = np.ascontiguousarray(x, np.float32).reshape(-1)
    q = np.empty(x.size, np.int32)
    s = C.c_float()
    rc = lib().ref_absmax_quant(_fp(x), C.c_int64(x.size), C.c_int(bits), _ip(q), C.byref(s))
    assert rc == 0
    return q, np.float32(s.value)
