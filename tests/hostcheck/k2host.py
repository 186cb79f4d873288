"""ctypes wrapper for the TEST-ONLY host compile of the K2 solve core."""
import ctypes as C

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build())
        _lib.k2_host_f16_to_f32.restype = C.c_float
        _lib.k2_host_f16_to_f32.argtypes = [C.c_uint16]
        _lib.k2_host_f32_to_f16.restype = C.c_uint16
        _lib.k2_host_f32_to_f16.argtypes = [C.c_float]
        _lib.k2_host_select_rank.restype = C.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def alloc_outputs(nt: int, stages: int):
    return dict(
        info=np.zeros(8, np.int32), sv=np.zeros(nt, np.float32), scal=np.zeros(4, np.float32),
        coef=np.zeros((nt, nt), np.float32), chigh=np.zeros((nt, nt), np.uint16),
        codes=np.zeros((nt, stages, nt), np.uint8), qscale=np.zeros((nt, stages), np.float32),
        qzp=np.zeros((nt, stages), np.float32), qres=np.zeros((nt, stages), np.float32),
        chat=np.zeros((nt, nt), np.float32), cbar=np.zeros(nt, np.float32), W=np.zeros((nt, nt), np.float32),
        gvec=np.zeros(nt, np.float32), V=np.zeros((nt, nt), np.float64))


def solve(G, dm, *, center=True, thr=0.9, max_rank=0, min_mask_size=10, bits=4, stages=2, has_mask=False,
          present=None, weights=None, avg_order=None, sign_ref=None):
    G = np.ascontiguousarray(G, np.float64)
    nt = G.shape[0]
    present = (1 << nt) - 1 if present is None else present
    weights = np.full(nt, 1.0 / nt) if weights is None else np.ascontiguousarray(weights, np.float64)
    avg_order = np.arange(nt, dtype=np.int32) if avg_order is None else np.ascontiguousarray(avg_order, np.int32)
    o = alloc_outputs(nt, stages)
    sr = None if sign_ref is None else np.ascontiguousarray(sign_ref, np.float64)
    lib().k2_host_solve(
        C.c_int(nt), C.c_int(int(center)), C.c_float(thr), C.c_int(max_rank or 0), C.c_int(min_mask_size),
        C.c_int(bits), C.c_int(stages), _p(G, C.c_double), C.c_int64(dm), C.c_int(int(has_mask)),
        C.c_uint32(present), _p(weights, C.c_double), _p(avg_order, C.c_int32),
        _p(sr, C.c_double) if sr is not None else None,
        _p(o["info"], C.c_int32), _p(o["sv"], C.c_float), _p(o["scal"], C.c_float), _p(o["coef"], C.c_float),
        _p(o["chigh"], C.c_uint16), _p(o["codes"], C.c_uint8), _p(o["qscale"], C.c_float), _p(o["qzp"], C.c_float),
        _p(o["qres"], C.c_float), _p(o["chat"], C.c_float), _p(o["cbar"], C.c_float), _p(o["W"], C.c_float),
        _p(o["gvec"], C.c_float), _p(o["V"], C.c_double))
    return o


def f32_to_f16(x: float) -> int:
    return int(lib().k2_host_f32_to_f16(C.c_float(x)))


def f16_to_f32(h: int) -> float:
    return float(lib().k2_host_f16_to_f32(C.c_uint16(h)))


def select_rank(S, thr, max_rank=0, min_rank=1):
    S = np.ascontiguousarray(S, np.float32)
    er = C.c_float()
    k = lib().k2_host_select_rank(_p(S, C.c_float), C.c_int(S.size), C.c_float(thr), C.c_int(max_rank or 0),
                                  C.c_int(min_rank), C.byref(er))
    return int(k), float(er.value)


def rtvq_short(x, bits, stages):
    x = np.ascontiguousarray(x, np.float32)
    n = x.size
    codes = np.zeros((stages, max(n, 1)), np.uint8)
    sc, zp, rn = (np.zeros(stages, np.float32) for _ in range(3))
    deq = np.zeros(max(n, 1), np.float32)
    lib().k2_host_rtvq_short(_p(x, C.c_float), C.c_int(n), C.c_int(bits), C.c_int(stages), _p(codes, C.c_uint8),
                             C.c_int(max(n, 1)), _p(sc, C.c_float), _p(zp, C.c_float), _p(rn, C.c_float),
                             _p(deq, C.c_float))
    return codes[:, :n], sc, zp, rn, deq[:n]
