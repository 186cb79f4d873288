"""Weighted averaging in coefficient space and reconstruction.  Mirror of src/svd_hybrid/merge.py:61-626.

These operator-level functions work on the reference's artifact structures (bases + compressed
coefficients) and serve the artifact reload path; the pipeline itself merges with the fused K3
kernel (engine.MergeJob), which folds averaging weights, reconstruction, mask scatter and
``base + delta`` into one pass."""
from typing import Dict, List, Optional, Tuple

import torch

from .. import _native
from .rtvq import RTVQQuantizer


def dequantize_and_average(compressed_coeffs: Dict[str, Dict], weights: Dict[str, float], quantizer: RTVQQuantizer,
                           region: str = "masked", device: str = "cpu") -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """merge.py:61-141: sorted task order; weights renormalised over the tasks that have the region."""
    names = sorted(compressed_coeffs.keys())
    hi, lo, w = [], [], []
    for n in names:
        art = compressed_coeffs[n]
        if art is None or art.get(region) is None:
            continue
        reg = art[region]
        hi.append(reg["c_high_fp16"].to(device).float())
        lo.append(quantizer.dequantize(reg["c_low_quant"], device=device).float())
        w.append(weights.get(n, 1.0 / len(names)))
    if not hi:
        return None, None
    tot = sum(w)
    wt = torch.tensor([x / tot for x in w], device=device, dtype=torch.float32).view(-1, 1)
    return (torch.stack(hi, 0) * wt).sum(0), (torch.stack(lo, 0) * wt).sum(0)


def reconstruct_from_coefficients(avg_c_high: torch.Tensor, avg_c_low: torch.Tensor, U_high: torch.Tensor,
                                  U_low: torch.Tensor, device: str = "cpu", mean: Optional[torch.Tensor] = None) -> torch.Tensor:
    """U_high c_high + U_low c_low (+ mean) in fp32 (merge.py:144-194), computed on the GPU."""
    _native.require_cuda()
    g = torch.device("cuda")
    out = U_high.to(g).float() @ avg_c_high.to(g).float() + U_low.to(g).float() @ avg_c_low.to(g).float()
    if mean is not None:
        out = out + mean.squeeze().to(g).float()
    return out.to(device)


def merge_parameter(param_name: str, compressed_params: Dict[str, Dict], basis: Dict, weights: Dict[str, float],
                    quantizer: RTVQQuantizer, original_shape: torch.Size, mask: Optional[torch.Tensor] = None,
                    include_noise: bool = False, noise_shrink: float = 1.0, device: str = "cpu") -> torch.Tensor:
    """merge.py:197-301."""
    from .mask_loader import reconstruct_from_masked

    def region(reg_basis, reg_name):
        if reg_basis is None:
            return None
        c_hi, c_lo = dequantize_and_average(compressed_params, weights, quantizer, region=reg_name, device=device)
        if c_hi is None:
            return None
        return reconstruct_from_coefficients(c_hi, c_lo, reg_basis["U_high"], reg_basis["U_low"], device=device,
                                             mean=reg_basis.get("mean"))

    merged_masked = region(basis.get("masked"), "masked")
    merged_unmasked = None
    if include_noise:
        merged_unmasked = region(basis.get("noise"), "unmasked")
        if merged_unmasked is not None:
            merged_unmasked = merged_unmasked * noise_shrink
    if mask is not None and merged_masked is not None:
        return reconstruct_from_masked(merged_masked, merged_unmasked, mask, original_shape)
    if merged_masked is not None:
        return merged_masked.view(original_shape)
    return torch.zeros(original_shape, device=device)


def merge_all_parameters(compressed_all: Dict[str, Dict[str, Dict]], bases: Dict[str, Dict], masks: Dict[str, torch.Tensor],
                         weights: Dict[str, float], original_shapes: Dict[str, torch.Size], config, device: str = "cpu",
                         verbose: bool = True) -> Dict[str, torch.Tensor]:
    """merge.py:304-426: every parameter of ``compressed_all`` in sorted order."""
    quantizer = RTVQQuantizer(num_bits=config.svd_low_bits, num_stages=config.svd_rtvq_stages)
    out = {}
    for name in sorted(compressed_all.keys()):
        out[name] = merge_parameter(name, compressed_all[name], bases[name], weights, quantizer, original_shapes[name],
                                    mask=masks.get(name), include_noise=config.svd_include_noise,
                                    noise_shrink=config.svd_noise_shrink, device=device)
    if verbose:
        total = sum(d.norm().item() ** 2 for d in out.values()) ** 0.5
        print(f"   merged {len(out)} parameters, total delta norm {total:.6f}")
    return out


def apply_merged_deltas(base_state_dict: Dict[str, torch.Tensor], merged_deltas: Dict[str, torch.Tensor],
                        device: str = "cpu", verbose: bool = True) -> Dict[str, torch.Tensor]:
    """merged[p] = base[p] + delta[p]; other keys are cloned (merge.py:429-552)."""
    out = {}
    for name, b in base_state_dict.items():
        out[name] = b + merged_deltas[name].to(b.device) if name in merged_deltas else b.clone()
    if verbose:
        print(f"   applied {sum(1 for n in base_state_dict if n in merged_deltas)} deltas to {len(out)} parameters")
    return out


def merge_with_clustering(compressed_all: Dict[str, Dict[str, Dict]], bases: Dict[str, Dict], masks: Dict[str, torch.Tensor],
                          weights: Dict[str, float], cluster_assignments: Dict[str, int],
                          original_shapes: Dict[str, torch.Size], config, device: str = "cpu") -> Dict[str, torch.Tensor]:
    """merge.py:555-626: merge inside each cluster, then softmax(mean member weight)-average the clusters."""
    from .clustering import get_cluster_members, merge_cluster_results
    clusters = get_cluster_members(cluster_assignments)
    per_cluster, score = {}, {}
    for cid, members in clusters.items():
        cw = {n: weights.get(n, 1.0) for n in members}
        tot = sum(cw.values())
        cw = {n: v / tot for n, v in cw.items()}
        sub = {p: {n: art[n] for n in members if n in art} for p, art in compressed_all.items()}
        per_cluster[cid] = merge_all_parameters(sub, bases, masks, cw, original_shapes, config, device, verbose=False)
        score[cid] = sum(weights.get(n, 1.0) for n in members) / len(members)
    return merge_cluster_results(per_cluster, score, device)
