"""Projection onto the bases and coefficient compression.  Mirror of src/svd_hybrid/compress.py:6-207.

Fine-grained operator API kept for parity tests and the artifact tools; ``compress_all_parameters``
on whole state dicts is served by the fused engine (closed-form coefficients in K2).  The explicit
projections here are tall-skinny matrix-vector products run by K14 (csrc/k14_operators.cu) on the device."""
from typing import Dict, Optional, Tuple

import torch

from . import _ops
from .rtvq import RTVQQuantizer


def project_to_basis(delta: torch.Tensor, U_high: torch.Tensor, U_low: torch.Tensor,
                     mean: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """c_high = U_high^T delta, c_low = U_low^T delta in fp32 (compress.py:6-21) on the device (K14,
    svdq_basis_project).  ``mean`` (extension) is subtracted inside the kernel.  Results on delta's device."""
    dev = delta.device
    return _ops.project(delta, U_high, mean).to(dev), _ops.project(delta, U_low, mean).to(dev)


def compress_single_task(task_delta: torch.Tensor, U_high: torch.Tensor, U_low: torch.Tensor, quantizer: RTVQQuantizer,
                         device: str = "cpu", mean: Optional[torch.Tensor] = None) -> Dict:
    """compress.py:24-56 -> {"c_high_fp16": fp16 CPU tensor, "c_low_quant": RTVQ object}."""
    c_high, c_low = project_to_basis(task_delta, U_high, U_low, mean=mean)
    c_high_fp16 = c_high if c_high.dtype == torch.float16 else c_high.half()
    return {"c_high_fp16": c_high_fp16.cpu(), "c_low_quant": quantizer.quantize(c_low.cpu())}


def compress_masked_regions(task_deltas_masked: Dict[str, torch.Tensor],
                            task_deltas_unmasked: Optional[Dict[str, torch.Tensor]], basis_masked: Optional[Dict],
                            basis_unmasked: Optional[Dict], quantizer: RTVQQuantizer, device: str = "cpu") -> Dict[str, Dict]:
    """compress.py:59-111 -> {task: {"masked": artifact | None, "unmasked": artifact | None}}."""
    out = {}
    for task, dm in task_deltas_masked.items():
        art = {"masked": None, "unmasked": None}
        if basis_masked is not None and len(dm) > 0:
            art["masked"] = compress_single_task(dm, basis_masked["U_high"], basis_masked["U_low"], quantizer, device,
                                                 mean=basis_masked.get("mean"))
        du = task_deltas_unmasked.get(task) if (basis_unmasked is not None and task_deltas_unmasked is not None) else None
        if du is not None and len(du) > 0:
            art["unmasked"] = compress_single_task(du, basis_unmasked["U_high"], basis_unmasked["U_low"], quantizer,
                                                   device, mean=basis_unmasked.get("mean"))
        out[task] = art
    return out


def compress_parameter(param_name: str, task_vectors: Dict[str, Dict[str, torch.Tensor]], mask: Optional[torch.Tensor],
                       basis: Dict, quantizer: RTVQQuantizer, include_noise: bool = False, min_mask_size: int = 10,
                       device: str = "cpu") -> Optional[Dict]:
    """compress.py:114-170."""
    from .mask_loader import apply_mask_to_tensor, get_unmasked_portion
    deltas = {t: tv[param_name] for t, tv in task_vectors.items() if param_name in tv}
    if not deltas:
        return None
    masked, unmasked = {}, {}
    for t, d in deltas.items():
        if mask is not None and mask.shape == d.shape:
            masked[t] = apply_mask_to_tensor(d, mask) if mask.sum() >= min_mask_size else torch.tensor([])
            if include_noise:
                unmasked[t] = get_unmasked_portion(d, mask)
        else:
            masked[t] = d.flatten()
    return compress_masked_regions(masked, unmasked if include_noise else None, basis.get("masked"),
                                   basis.get("noise") if include_noise else None, quantizer, device)


def compress_all_parameters(task_vectors: Dict[str, Dict[str, torch.Tensor]], masks: Dict[str, torch.Tensor],
                            bases: Dict[str, Dict], config, device: str = "cpu") -> Dict[str, Dict]:
    """compress.py:173-207 (operator-by-operator; the pipeline uses engine.merge_state_dicts instead)."""
    quantizer = RTVQQuantizer(num_bits=config.svd_low_bits, num_stages=config.svd_rtvq_stages)
    out = {}
    for name in sorted(bases.keys()):
        c = compress_parameter(name, task_vectors, masks.get(name), bases[name], quantizer,
                               include_noise=config.svd_include_noise, min_mask_size=config.svd_min_mask_size,
                               device=device)
        if c is not None:
            out[name] = c
    return out
