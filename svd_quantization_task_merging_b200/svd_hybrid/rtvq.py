"""RTVQ: multi-stage residual quantisation.  Mirror of src/svd_hybrid/rtvq.py:4-161.

Same names, argument order, defaults and return structures as the reference; the arithmetic runs in
the K4 kernels of libsvdq.so (bit-exact fp32 operation order: reciprocal-then-multiply scale,
half-to-even rounding, separate multiply and add, no guards -- NaN/inf propagate exactly like the
reference when max == min).  Inputs may live on any device; they are moved to the GPU, and results
are returned where the reference returns them.
"""
from typing import Dict, List, Tuple

import torch

from .. import _native


def _gpu_f32(x: torch.Tensor) -> torch.Tensor:
    _native.require_cuda()
    y = x.detach().to(device="cuda", dtype=torch.float32).contiguous().view(-1)
    if y.data_ptr() % 16:
        y = y.clone()
    return y


def _quantize_device(x: torch.Tensor, bits: int, stages: int, code_bytes: int, packed: bool = False):
    """x: flat fp32 CUDA tensor -> (codes[stages, ld], scale[stages], zp[stages], resnorm[stages][, packed words])."""
    n = x.numel()
    ld = (n + 15) // 16 * 16
    cdt = torch.uint8 if code_bytes == 1 else torch.int16
    codes = torch.zeros(stages, ld, dtype=cdt, device=x.device)
    scale = torch.zeros(stages, dtype=torch.float32, device=x.device)
    zp = torch.zeros_like(scale)
    rn = torch.zeros_like(scale)
    scratch = torch.empty(_native.load().svdq_k4_scratch_bytes(), dtype=torch.uint8, device=x.device)
    pk, pld = None, 0
    if packed:
        pld = (n * bits + 31) // 32 + 8
        pk = torch.zeros(stages, pld, dtype=torch.int32, device=x.device)
    with torch.cuda.device(x.device):
        _native.call("svdq_rtvq_quantize", x.data_ptr(), n, bits, stages, codes.data_ptr(), ld, code_bytes,
                     scale.data_ptr(), zp.data_ptr(), rn.data_ptr(), scratch.data_ptr(),
                     pk.data_ptr() if pk is not None else None, pld, _native.stream_ptr())
    if packed:
        return codes, scale, zp, rn, pk
    return codes, scale, zp, rn


def asymmetric_quantization(X: torch.Tensor, qbit: int = 8, verbose: bool = False
                            ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """rtvq.py:4-27.  -> (X_q uint8 | int16, scale 0-d fp32, zero_point 0-d fp32) on X's device."""
    if X.numel() == 0:
        raise RuntimeError("min(): Expected reduction dim to be specified for input.numel() == 0")
    if not (qbit <= 8 or qbit == 16):
        raise ValueError(f"qbit must be <= 8 or == 16, got {qbit}")
    x = _gpu_f32(X)
    codes, scale, zp, _ = _quantize_device(x, qbit, 1, 1 if qbit <= 8 else 2)
    q = codes[0, : x.numel()].view(X.shape)
    return q.to(X.device), scale[0].to(X.device), zp[0].to(X.device)


def asymmetric_dequantization(quantized: torch.Tensor, scale: torch.Tensor, zero_point: torch.Tensor) -> torch.Tensor:
    """rtvq.py:29-36: (q - zero_point) / scale, on the device of ``quantized``."""
    _native.require_cuda()
    if quantized.numel() == 0:
        return torch.zeros(quantized.shape, dtype=torch.float32, device=quantized.device)
    cb = 2 if quantized.dtype == torch.int16 else 1
    q = quantized.detach().to("cuda").contiguous().view(-1)
    if cb == 1 and q.dtype != torch.uint8:
        q = q.to(torch.uint8)
    sc = torch.as_tensor(scale, dtype=torch.float32).reshape(1).to("cuda")
    zp = torch.as_tensor(zero_point, dtype=torch.float32).reshape(1).to("cuda")
    out = torch.empty(q.numel(), dtype=torch.float32, device="cuda")
    _native.call("svdq_rtvq_dequantize", q.data_ptr(), q.numel(), cb, 1, q.numel(), sc.data_ptr(), zp.data_ptr(),
                 out.data_ptr(), _native.stream_ptr())
    return out.view(quantized.shape).to(quantized.device)


def multistage_residual_quantization(tensor: torch.Tensor, num_bits: int = 4, num_stages: int = 2,
                                     verbose: bool = False, packed: bool = False) -> List[Dict]:
    """rtvq.py:39-82.  Payload tensors are returned on the CPU like the reference.  ``packed=True`` (an
    addition, bits in {1,2,4,8}) adds a "packed" int32 tensor per stage: the same codes, ``num_bits`` each."""
    if tensor.numel() == 0:
        return []
    x = _gpu_f32(tensor)
    out = _quantize_device(x, num_bits, num_stages, 1, packed=packed)
    codes, scale, zp, rn = out[:4]
    codes_h = codes[:, : x.numel()].cpu()
    scale_h, zp_h, rn_h = scale.cpu(), zp.cpu(), rn.cpu()
    pay = [{"stage": s, "quantized": codes_h[s].clone().view(tensor.shape), "scale": scale_h[s].clone(),
            "zero_point": zp_h[s].clone(), "residual_norm": rn_h[s].item()} for s in range(num_stages)]
    if packed:
        words = (x.numel() * num_bits + 31) // 32
        pk = out[4][:, :words].cpu()
        for s in range(num_stages):
            pay[s]["packed"] = pk[s].clone()
    return pay


def multistage_residual_dequantization(payloads: List[Dict], device: str = "cpu") -> torch.Tensor:
    """rtvq.py:85-103: left-to-right sum of the stage dequantisations."""
    if not payloads:
        return torch.tensor([], device=device)
    _native.require_cuda()
    shape = payloads[0]["quantized"].shape
    n = payloads[0]["quantized"].numel()
    S = len(payloads)
    ld = (n + 15) // 16 * 16
    codes = torch.zeros(S, ld, dtype=torch.uint8, device="cuda")
    for s, p in enumerate(payloads):
        codes[s, :n] = p["quantized"].reshape(-1).to("cuda")
    sc = torch.stack([torch.as_tensor(p["scale"], dtype=torch.float32).reshape(()) for p in payloads]).to("cuda")
    zp = torch.stack([torch.as_tensor(p["zero_point"], dtype=torch.float32).reshape(()) for p in payloads]).to("cuda")
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    _native.call("svdq_rtvq_dequantize", codes.data_ptr(), ld, 1, S, n, sc.data_ptr(), zp.data_ptr(),
                 out.data_ptr(), _native.stream_ptr())
    return out.view(shape).to(device)


class RTVQQuantizer:
    """rtvq.py:106-139."""

    def __init__(self, num_bits: int = 4, num_stages: int = 2):
        self.num_bits = num_bits
        self.num_stages = num_stages

    def quantize(self, tensor: torch.Tensor) -> Dict:
        return {"payloads": multistage_residual_quantization(tensor, self.num_bits, self.num_stages),
                "num_bits": self.num_bits, "num_stages": self.num_stages,
                "original_shape": tensor.shape, "original_dtype": str(tensor.dtype)}

    def dequantize(self, quantized_obj: Dict, device: str = "cpu") -> torch.Tensor:
        out = multistage_residual_dequantization(quantized_obj["payloads"], device=device)
        if "original_shape" in quantized_obj:
            out = out.view(quantized_obj["original_shape"])
        return out


def estimate_compression_ratio(original: torch.Tensor, quantized_obj: Dict) -> float:
    """rtvq.py:142-161 (byte accounting on the low-energy block only)."""
    n = original.numel()
    stages, bits = quantized_obj["num_stages"], quantized_obj["num_bits"]
    return (n * 4) / max(n * bits / 8 * stages + 8 * stages, 1)
