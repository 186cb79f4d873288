"""Tall-mask loading, combination and application.  Mirror of src/svd_hybrid/mask_loader.py:66-763.

Mask combination runs in the ``k_combine_masks`` kernel (libsvdq.so); the fused path combines masks
inside K1 and never materialises the torch.bool result.  Gather / scatter through a mask are plain
torch indexing (data movement), executed on the GPU."""
import copy
import os
from collections import OrderedDict
from typing import Dict, List, Optional

import numpy as np
import torch

from .. import _native
from . import _ops

_STRATEGIES = ("union", "intersection", "majority")


def state_dict_to_vector(state_dict: Dict[str, torch.Tensor], remove_keys: Optional[List[str]] = None) -> torch.Tensor:
    """Flatten in sorted-key order (mask_loader.py:66-105)."""
    skip = set(remove_keys or [])
    parts = [state_dict[k].flatten() for k in sorted(state_dict) if k not in skip]
    return torch.cat(parts) if parts else torch.tensor([])


def vector_to_state_dict(vector: torch.Tensor, reference_state_dict: Dict[str, torch.Tensor],
                         remove_keys: Optional[List[str]] = None) -> "OrderedDict[str, torch.Tensor]":
    """Inverse of state_dict_to_vector.  The reference's version raises NameError (it reads an undefined
    ``state_dict``, mask_loader.py:110); this is the behaviour its own tests expect
    (tests/test_mask_strategies.py:262-330)."""
    skip = set(remove_keys or [])
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    off = 0
    for k in sorted(reference_state_dict):
        if k in skip:
            continue
        ref = reference_state_dict[k]
        n = ref.numel()
        out[k] = vector[off: off + n].view(ref.shape).clone()
        off += n
    if "transformer.shared.weight" in out:
        for k in skip:
            out[k] = out["transformer.shared.weight"]
    return out


def load_tall_mask_file(mask_path: str, reference_state_dict: Dict[str, torch.Tensor],
                        remove_keys: Optional[List[str]] = None, device: str = "cpu") -> Dict[str, Dict[str, torch.Tensor]]:
    """TALL mask container: {task: bit-packed uint8 array}, np.unpackbits order (mask_loader.py:125-206)."""
    if not os.path.exists(mask_path):
        raise FileNotFoundError(f"TALL mask file not found: {mask_path}")
    remove_keys = remove_keys or []
    packed_masks = torch.load(mask_path, map_location=device, weights_only=False)
    n_expected = sum(v.numel() for k, v in reference_state_dict.items() if k not in remove_keys)
    out = {}
    for task, packed in packed_masks.items():
        if isinstance(packed, torch.Tensor):
            packed = packed.cpu().numpy()
        elif not isinstance(packed, np.ndarray):
            raise TypeError(f"Unexpected type for packed_mask: {type(packed)}")
        vec = torch.from_numpy(np.unpackbits(packed)).to(device)[:n_expected]
        sd = vector_to_state_dict(vec, reference_state_dict, remove_keys=remove_keys)
        out[task] = {k: v.bool() for k, v in sd.items()}
    return out


def load_single_mask(mask_path: str, device: str = "cpu") -> Dict[str, torch.Tensor]:
    if not os.path.exists(mask_path):
        raise FileNotFoundError(f"Mask file not found: {mask_path}")
    masks = torch.load(mask_path, map_location=device, weights_only=False)
    for k, m in masks.items():
        if m.dtype != torch.bool:
            masks[k] = m.bool()
    return masks


def load_task_masks(mask_dir: str, task_names: List[str], device: str = "cpu",
                    reference_state_dict: Optional[Dict[str, torch.Tensor]] = None,
                    remove_keys: Optional[List[str]] = None, verbose: bool = True) -> Dict[str, Optional[Dict[str, torch.Tensor]]]:
    """mask_loader.py:242-409: TALL_mask_{n}task[s].npy container first, else {t}_mask.pt | {t}.pt | {t}/mask.pt."""
    n = len(task_names)
    tall = next((p for p in (os.path.join(mask_dir, f"{pre}_mask_{n}{suf}.npy") for pre in ("TALL", "tall")
                             for suf in ("task", "tasks")) if os.path.exists(p)), None)
    if tall is not None and reference_state_dict is not None:
        everything = load_tall_mask_file(tall, reference_state_dict, remove_keys=remove_keys, device=device)
        return {t: everything.get(t) for t in task_names}
    out: Dict[str, Optional[Dict[str, torch.Tensor]]] = {}
    for t in task_names:
        cands = [os.path.join(mask_dir, f"{t}_mask.pt"), os.path.join(mask_dir, f"{t}.pt"),
                 os.path.join(mask_dir, t, "mask.pt")]
        hit = next((c for c in cands if os.path.exists(c)), None)
        out[t] = load_single_mask(hit, device) if hit is not None else None      # None: "use all parameters"
        if hit is None and verbose:
            print(f"   no mask found for {t}: all parameters are used")
    return out


def _combine(masks: List[torch.Tensor], strategy: str) -> torch.Tensor:
    if not masks:
        raise ValueError("Empty mask list")
    if strategy not in _STRATEGIES:
        raise ValueError(f"Unknown mask strategy: {strategy}")
    _native.require_cuda()
    dev_in, shape = masks[0].device, masks[0].shape
    flat = []
    for m in masks:
        m = m.detach()
        if m.dtype != torch.bool:
            m = m.bool()
        flat.append(m.to("cuda").contiguous().view(-1))
    n = flat[0].numel()
    ptrs = torch.tensor([m.data_ptr() for m in flat], dtype=torch.int64, device="cuda")
    out = torch.empty(n, dtype=torch.bool, device="cuda")
    _native.call("svdq_combine_masks", ptrs.data_ptr(), len(flat), n, _native.STRATEGY_CODE[strategy], out.data_ptr(),
                 _native.stream_ptr())
    return out.view(shape).to(dev_in)


def compute_union_mask(masks: List[torch.Tensor]) -> torch.Tensor:
    return _combine(masks, "union")


def compute_intersection_mask(masks: List[torch.Tensor]) -> torch.Tensor:
    return _combine(masks, "intersection")


def compute_majority_mask(masks: List[torch.Tensor], threshold: float = 0.5) -> torch.Tensor:
    """votes >= threshold * n (mask_loader.py:456-485).  The kernel implements the reference's default
    threshold 0.5 exactly (2 * votes >= n); other thresholds are not used anywhere in the reference."""
    if threshold != 0.5:
        raise ValueError("only the reference's default threshold 0.5 is supported")
    return _combine(masks, "majority")


def combine_masks(task_masks: Dict[str, Optional[Dict[str, torch.Tensor]]], strategy: str = "union",
                  device: str = "cpu", verbose: bool = True) -> Dict[str, torch.Tensor]:
    """mask_loader.py:488-648: per parameter, combine the masks of the tasks that have one."""
    if not task_masks:
        return {}
    names = set()
    for pm in task_masks.values():
        if pm is not None:
            names.update(pm.keys())
    if names and strategy not in _STRATEGIES:
        raise ValueError(f"Unknown mask strategy: {strategy}")
    out = {}
    for name in names:
        present = [pm[name] for pm in task_masks.values() if pm is not None and name in pm]
        if present:
            out[name] = _combine(present, strategy).to(device)
    return out


def apply_mask_to_tensor(tensor: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """tensor.flatten()[mask.flatten()] (mask_loader.py:665-686) by the device selection kernels (K14)."""
    if tensor.shape != mask.shape:
        raise ValueError(f"Shape mismatch: tensor {tensor.shape} vs mask {mask.shape}")
    return _ops.mask_select(tensor, mask)


def get_unmasked_portion(tensor: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """tensor.flatten()[~mask.flatten()] (mask_loader.py:689-709)."""
    if tensor.shape != mask.shape:
        raise ValueError(f"Shape mismatch: tensor {tensor.shape} vs mask {mask.shape}")
    return _ops.mask_select(tensor, mask, invert=True)


def reconstruct_from_masked(masked_values: torch.Tensor, unmasked_values: Optional[torch.Tensor], mask: torch.Tensor,
                            original_shape: torch.Size) -> torch.Tensor:
    """Scatter back through the mask, zeros elsewhere unless a noise part is given (mask_loader.py:712-763).
    Result on masked_values' device."""
    home = masked_values.device
    dev = home if home.type == "cuda" else torch.device("cuda")
    _native.require_cuda()
    out = torch.zeros(mask.numel(), dtype=masked_values.dtype, device=dev)
    _ops.mask_scatter(masked_values, mask, out)
    if unmasked_values is not None:
        _ops.mask_scatter(unmasked_values, mask, out, invert=True)
    return out.view(original_shape).to(home)
