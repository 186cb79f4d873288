"""Device operators behind the fine-grained mirrors (K14, csrc/k14_operators.cu): projection on a stored basis,
expansion of averaged coefficients, selection / scatter through a boolean mask.  Thin ctypes calls on torch device
memory -- no torch compute (matmul, boolean indexing) on these paths, and no CPU fallback."""
from typing import Optional, Tuple

import torch

from .. import _native

_ELEM_OK = (1, 2, 4, 8)


def _cuda(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_cuda else t.to("cuda")


def _basis(U: torch.Tensor) -> Tuple[torch.Tensor, bool]:
    """-> (2-D row-major device tensor in fp16 or fp32, is_fp16)."""
    U = _cuda(U)
    if U.dim() == 1:
        U = U.view(-1, 1)
    if U.dtype not in (torch.float16, torch.float32):
        U = U.float()
    if U.shape[1] > 0 and U.stride(1) != 1:
        U = U.contiguous()
    if U.shape[0] > 1 and U.stride(0) < U.shape[1]:
        U = U.contiguous()
    return U, U.dtype == torch.float16


def _ld(U: torch.Tensor) -> int:
    return U.stride(0) if U.shape[0] > 1 else max(U.shape[1], 1)


def _f32(v: torch.Tensor) -> torch.Tensor:
    return _cuda(v).reshape(-1).float().contiguous()


def project(delta: torch.Tensor, U: torch.Tensor, mean: Optional[torch.Tensor] = None) -> torch.Tensor:
    """U^T (delta - mean) -> fp32 [cols] on the device (svdq_basis_project, 32 columns per launch)."""
    _native.require_cuda()
    U, f16 = _basis(U)
    rows, cols = U.shape
    d = _f32(delta)
    if d.numel() != rows:
        raise ValueError(f"size mismatch: basis has {rows} rows, vector has {d.numel()} elements")
    m = _f32(mean) if mean is not None else None
    if m is not None and m.numel() != rows:
        raise ValueError(f"size mismatch: basis has {rows} rows, mean has {m.numel()} elements")
    out = torch.zeros(cols, dtype=torch.float32, device=U.device)
    if cols == 0 or rows == 0:
        return out
    scratch = torch.empty(_native.load().svdq_project_scratch_bytes(), dtype=torch.uint8, device=U.device)
    es = U.element_size()
    with torch.cuda.device(U.device):
        for c0 in range(0, cols, 32):
            n = min(32, cols - c0)
            _native.call("svdq_basis_project", int(f16), U.data_ptr() + c0 * es, _ld(U), n, rows, d.data_ptr(),
                         m.data_ptr() if m is not None else None, out.data_ptr() + 4 * c0, scratch.data_ptr(),
                         _native.stream_ptr())
    return out


def expand(c_high: torch.Tensor, c_low: Optional[torch.Tensor], U_high: torch.Tensor, U_low: Optional[torch.Tensor],
           mean: Optional[torch.Tensor] = None, scale: float = 1.0) -> torch.Tensor:
    """scale * (U_high c_high + U_low c_low + mean) -> fp32 [rows] on the device (svdq_basis_expand)."""
    _native.require_cuda()
    Uh, f16 = _basis(U_high)
    rows, k = Uh.shape
    Ul, nl = None, 0
    if U_low is not None and U_low.numel() > 0:
        Ul, f16l = _basis(U_low)
        if f16l != f16:
            Uh, Ul, f16 = Uh.float(), Ul.float(), False
        if Ul.shape[0] != rows:
            raise ValueError(f"size mismatch: U_high has {rows} rows, U_low has {Ul.shape[0]}")
        nl = Ul.shape[1]
    ch = _f32(c_high)
    cl = _f32(c_low) if nl else None
    if ch.numel() != k or (nl and cl.numel() != nl):
        raise ValueError("size mismatch between coefficients and basis columns")
    m = _f32(mean) if mean is not None else None
    if m is not None and m.numel() != rows:
        raise ValueError(f"size mismatch: basis has {rows} rows, mean has {m.numel()} elements")
    out = torch.empty(rows, dtype=torch.float32, device=Uh.device)
    with torch.cuda.device(Uh.device):
        _native.call("svdq_basis_expand", int(f16), Uh.data_ptr(), _ld(Uh), k, Ul.data_ptr() if nl else None,
                     _ld(Ul) if nl else 0, nl, rows, ch.data_ptr(), cl.data_ptr() if nl else None,
                     m.data_ptr() if m is not None else None, float(scale), out.data_ptr(), _native.stream_ptr())
    return out


def _mask_bytes(mask: torch.Tensor, device) -> torch.Tensor:
    m = mask.to(device).reshape(-1)
    if m.dtype != torch.bool:
        m = m != 0
    return m.contiguous().view(torch.uint8)


def _offsets(mb: torch.Tensor, invert: bool) -> Tuple[torch.Tensor, int]:
    """-> (per-chunk offsets on the device, number of kept elements).  The count is read back (the output size of a
    selection is data dependent, exactly as with torch's boolean indexing)."""
    n = mb.numel()
    chunk = _native.load().svdq_select_chunk_elems()
    n_chunks = (n + chunk - 1) // chunk
    off = torch.empty(n_chunks + 1, dtype=torch.int64, device=mb.device)
    _native.call("svdq_mask_offsets", mb.data_ptr() if n else None, n, int(invert), off.data_ptr(), _native.stream_ptr())
    return off, int(off[-1].item())


def mask_select(tensor: torch.Tensor, mask: torch.Tensor, invert: bool = False) -> torch.Tensor:
    """tensor.flatten()[mask.flatten()] (or [~mask] with invert) on the device; result on tensor's device."""
    _native.require_cuda()
    home = tensor.device
    x = _cuda(tensor).reshape(-1).contiguous()
    if x.element_size() not in _ELEM_OK:
        raise ValueError(f"unsupported element size {x.element_size()}")
    with torch.cuda.device(x.device):
        mb = _mask_bytes(mask, x.device)
        off, kept = _offsets(mb, invert)
        out = torch.empty(kept, dtype=x.dtype, device=x.device)
        if kept:
            _native.call("svdq_mask_select", x.data_ptr(), x.element_size(), mb.data_ptr(), x.numel(), int(invert),
                         off.data_ptr(), out.data_ptr(), _native.stream_ptr())
    return out.to(home)


def mask_scatter(values: torch.Tensor, mask: torch.Tensor, out_flat: torch.Tensor, invert: bool = False) -> None:
    """out_flat[mask] = values (or out_flat[~mask] with invert), in place on a flat contiguous device tensor."""
    _native.require_cuda()
    if not out_flat.is_cuda or not out_flat.is_contiguous() or out_flat.dim() != 1:
        raise ValueError("out_flat must be a flat contiguous CUDA tensor")
    v = values.to(out_flat.device, out_flat.dtype).reshape(-1).contiguous()
    with torch.cuda.device(out_flat.device):
        mb = _mask_bytes(mask, out_flat.device)
        if mb.numel() != out_flat.numel():
            raise ValueError(f"size mismatch: mask has {mb.numel()} elements, output has {out_flat.numel()}")
        off, kept = _offsets(mb, invert)
        if kept != v.numel():
            raise ValueError(f"shape mismatch: {v.numel()} values for {kept} selected positions")
        if kept:
            _native.call("svdq_mask_scatter", v.data_ptr(), v.element_size(), mb.data_ptr(), mb.numel(), int(invert),
                         off.data_ptr(), out_flat.data_ptr(), _native.stream_ptr())
