"""Whole-tensor TVQ quantisers.  Mirror of the reference's root ``quantization_utils.py:60-211``
(same names -- including the ``qunatization_error_check`` spelling -- and the same arithmetic,
``dequantize_absmax`` multiplying by the scale included); compute runs in the K4 kernels."""
from typing import Tuple

import torch

from . import _native
from .svd_hybrid.rtvq import asymmetric_dequantization as _asym_dequant
from .svd_hybrid.rtvq import asymmetric_quantization as _asym_quant

FLOAT32_BITS = 32


def absmax_quantization(X: torch.Tensor, qbit: int = 8, verbose: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """quantization_utils.py:60-73: s = (2^(b-1)-1)/max|X|; X_q = round(s X) as int8 / int16 (no clamp)."""
    _native.require_cuda()
    if X.numel() == 0:
        raise RuntimeError("max(): Expected reduction dim to be specified for input.numel() == 0")
    if not (qbit <= 8 or qbit == 16):
        raise ValueError(f"qbit must be <= 8 or == 16, got {qbit}")
    x = X.detach().to(device="cuda", dtype=torch.float32).contiguous().view(-1)
    cb = 1 if qbit <= 8 else 2
    q = torch.empty(x.numel(), dtype=torch.int8 if cb == 1 else torch.int16, device="cuda")
    s = torch.zeros(1, dtype=torch.float32, device="cuda")
    scratch = torch.empty(_native.load().svdq_k4_scratch_bytes(), dtype=torch.uint8, device="cuda")
    _native.call("svdq_absmax_quantize", x.data_ptr(), x.numel(), qbit, q.data_ptr(), cb, s.data_ptr(),
                 scratch.data_ptr(), _native.stream_ptr())
    return q.view(X.shape).to(X.device), s[0].to(X.device)


def asymmetric_quantization(X: torch.Tensor, qbit: int = 8, verbose: bool = False):
    """quantization_utils.py:76-99 (identical arithmetic to src/svd_hybrid/rtvq.py:4-27)."""
    return _asym_quant(X, qbit, verbose)


def dequantize_absmax(X_q: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    """quantization_utils.py:102-134: X_q.float() * scale (sic -- the reference multiplies)."""
    return X_q.float() * scale


def dequantize_asymmetric(X_q: torch.Tensor, scale: torch.Tensor, zero_point: torch.Tensor) -> torch.Tensor:
    """quantization_utils.py:137-172."""
    return _asym_dequant(X_q, scale, zero_point)


def _accumulated_abs_error(original_state_dict, quantized_state_dict, code_dtype, decode) -> torch.Tensor:
    total = 0
    for key, orig in original_state_dict.items():
        q = quantized_state_dict[key]
        rec = decode(key, q) if q.dtype == code_dtype else q
        total = total + torch.sum(torch.abs(orig - rec))
    return total


def qunatization_error_check(original_state_dict, quantized_state_dict):
    """quantization_utils.py:175-191: prints the summed |error|; int8 entries are decoded as q / scale."""
    def decode(key, q):
        return q.to(torch.float) / quantized_state_dict[key + "_qscale"]
    print(f"accumuated Quantized error: {_accumulated_abs_error(original_state_dict, quantized_state_dict, torch.int8, decode)}")


def quantization_error_check_asymmetric(original_state_dict, quantized_state_dict):
    """quantization_utils.py:194-211: uint8 entries are decoded as (q - zero_point) / scale."""
    def decode(key, q):
        zp = quantized_state_dict[key + "_qzeropoint"].to(torch.float)
        return (q.to(torch.float) - zp) / quantized_state_dict[key + "_qscale"]
    print(f"accumuated Quantized error: {_accumulated_abs_error(original_state_dict, quantized_state_dict, torch.uint8, decode)}")
