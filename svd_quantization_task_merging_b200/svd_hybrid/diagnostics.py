"""Reconstruction diagnostics and compression accounting.  Mirror of src/svd_hybrid/diagnostics.py:72-746.

The pipeline obtains the per-task error figures from reductions fused into the K3 kernel
(results.build_diagnostics); the functions here recompute them from artifact structures."""
import math
from typing import Dict, Optional

import numpy as np
import torch

from .. import _native
from . import _ops
from .rtvq import RTVQQuantizer, estimate_compression_ratio


def compute_reconstruction_error(original_delta: torch.Tensor, reconstructed_delta: torch.Tensor) -> Dict[str, float]:
    """diagnostics.py:72-117."""
    err = original_delta - reconstructed_delta
    on, en = original_delta.norm().item(), err.norm().item()
    return {"absolute_error": en, "relative_error": en / on if on > 1e-10 else 0,
            "max_absolute_error": err.abs().max().item(), "mean_absolute_error": err.abs().mean().item(),
            "original_norm": on, "reconstructed_norm": reconstructed_delta.norm().item()}


def compute_parameter_diagnostics(param_name: str, task_vectors: Dict[str, Dict[str, torch.Tensor]],
                                  compressed_params: Dict[str, Dict], basis: Dict, mask: Optional[torch.Tensor],
                                  quantizer: RTVQQuantizer, device: str = "cpu") -> Dict:
    """diagnostics.py:120-231.  The reconstruction deliberately omits the mean, as the reference does."""
    from .mask_loader import apply_mask_to_tensor
    d = {"param_name": param_name, "original_shape": None, "masked_size": 0, "unmasked_size": 0,
         "reconstruction_errors": {}, "compression_ratios": {}}
    bm = basis.get("masked")
    if bm is None:
        return d
    _native.require_cuda()
    g = torch.device("cuda")
    first = next(iter(task_vectors.keys()))
    if param_name in task_vectors[first]:
        d["original_shape"] = list(task_vectors[first][param_name].shape)
    if mask is not None:
        d["masked_size"] = int(mask.sum().item())
        d["unmasked_size"] = int((~mask).sum().item())
    else:
        d["masked_size"] = np.prod(d["original_shape"])
    d["basis"] = {"k": bm["k"], "D": bm["D"], "N": bm["N"], "energy_retained": bm["energy_retained"]}
    rel = []
    for task, tv in task_vectors.items():
        if param_name not in tv or task not in compressed_params:
            continue
        art = compressed_params[task]
        if art.get("masked") is None:
            continue
        orig = tv[param_name]
        orig = apply_mask_to_tensor(orig, mask.to(orig.device)) if (mask is not None and mask.shape == orig.shape) else orig.flatten()
        c_hi = art["masked"]["c_high_fp16"].to(g).float()
        c_lo_obj = art["masked"]["c_low_quant"]
        c_lo = quantizer.dequantize(c_lo_obj, device="cuda").float()
        rec = _ops.expand(c_hi, c_lo, bm["U_high"], bm["U_low"])
        em = compute_reconstruction_error(orig.to(g).float(), rec)
        rel.append(em["relative_error"])
        d["reconstruction_errors"][task] = em
        d["compression_ratios"][task] = estimate_compression_ratio(c_lo, c_lo_obj)
    if rel:
        d["mean_relative_error"] = float(np.mean(rel))
        d["std_relative_error"] = float(np.std(rel))
        d["max_relative_error"] = float(np.max(rel))
        d["min_relative_error"] = float(np.min(rel))
    return d


def summarize(per_parameter: Dict[str, Dict]) -> Dict:
    ranks, energy, errs, ratios = [], [], [], []
    for d in per_parameter.values():
        if "basis" in d:
            ranks.append(d["basis"]["k"])
            energy.append(d["basis"]["energy_retained"])
        if "mean_relative_error" in d:
            errs.append(d["mean_relative_error"])
        if d.get("compression_ratios"):
            ratios.append(np.mean(list(d["compression_ratios"].values())))
    return {"num_parameters": len(per_parameter),
            "average_rank": float(np.mean(ranks)) if ranks else 0, "std_rank": float(np.std(ranks)) if ranks else 0,
            "average_energy_retained": float(np.mean(energy)) if energy else 0,
            "average_reconstruction_error": float(np.mean(errs)) if errs else 0,
            "average_compression_ratio": float(np.mean(ratios)) if ratios else 0}


def compute_all_diagnostics(task_vectors, compressed_all, bases, masks, config, device: str = "cpu") -> Dict:
    """diagnostics.py:234-321 -> {"config", "per_parameter", "summary"}."""
    quantizer = RTVQQuantizer(num_bits=config.svd_low_bits, num_stages=config.svd_rtvq_stages)
    out = {"config": {"svd_energy_threshold": config.svd_energy_threshold, "svd_max_rank": config.svd_max_rank,
                      "svd_low_bits": config.svd_low_bits, "svd_rtvq_stages": config.svd_rtvq_stages,
                      "svd_mask_strategy": config.svd_mask_strategy, "svd_weighting": config.svd_weighting},
           "per_parameter": {}, "summary": {}}
    for name in sorted(bases.keys()):
        if name in compressed_all:
            out["per_parameter"][name] = compute_parameter_diagnostics(name, task_vectors, compressed_all[name],
                                                                       bases[name], masks.get(name), quantizer, device)
    out["summary"] = summarize(out["per_parameter"])
    return out


def compression_statistics_from_sizes(original_numel: Dict[str, int], n_tasks_with_param: Dict[str, int],
                                      sizes: Dict[str, Dict], num_tasks: int, num_bits: int, num_stages: int,
                                      task_names=None) -> Dict:
    """Byte accounting of diagnostics.py:385-566 from sizes alone: originals numel*4 B; fp16 c_high k*2 B per
    task; RTVQ ceil(n*b/8)+8 B per stage (none for an empty low block); bases (|U_high|+|U_low|)*2 B.
    ``sizes[param]`` = {"k", "r", "D", "n_tasks"}."""
    stats = {"original": {"total_bytes": 0, "per_task_bytes": {}, "per_param_bytes": {}},
             "compressed": {"total_bytes": 0, "fp16_high_energy_bytes": 0, "rtvq_low_energy_bytes": 0,
                            "svd_bases_bytes": 0, "per_task_bytes": {}, "per_param_bytes": {}},
             "per_parameter": {}, "summary": {}}
    for p, n in original_numel.items():
        stats["original"]["per_param_bytes"][p] = n * 4 * n_tasks_with_param.get(p, num_tasks)
    stats["original"]["total_bytes"] = sum(stats["original"]["per_param_bytes"].values())
    if task_names:
        per_task = sum(n * 4 for n in original_numel.values())
        stats["original"]["per_task_bytes"] = {t: per_task for t in task_names}
    tot_fp16 = tot_rtvq = tot_bases = 0
    for p, s in sizes.items():
        k, r, D, nt = s["k"], s["r"], s["D"], s["n_tasks"]
        n_low = r - k
        fp16_b = nt * k * 2
        rtvq_b = nt * num_stages * (math.ceil(n_low * num_bits / 8) + 8) if n_low > 0 else 0
        bases_b = (D * k + D * n_low) * 2
        comp = fp16_b + rtvq_b + bases_b
        orig = stats["original"]["per_param_bytes"].get(p, 0)
        stats["per_parameter"][p] = {"original_bytes": orig, "compressed_bytes": comp, "fp16_high_energy_bytes": fp16_b,
                                     "rtvq_low_energy_bytes": rtvq_b, "svd_bases_bytes": bases_b, "k": k, "D": D,
                                     "compression_ratio": orig / max(comp, 1) if orig > 0 else 0}
        stats["compressed"]["per_param_bytes"][p] = comp
        tot_fp16 += fp16_b
        tot_rtvq += rtvq_b
        tot_bases += bases_b
    ct = tot_fp16 + tot_rtvq + tot_bases
    stats["compressed"].update(fp16_high_energy_bytes=tot_fp16, rtvq_low_energy_bytes=tot_rtvq,
                               svd_bases_bytes=tot_bases, total_bytes=ct)
    ot = stats["original"]["total_bytes"]
    stats["summary"] = {"original_size_mb": ot / (1024 * 1024), "compressed_size_mb": ct / (1024 * 1024),
                        "overall_compression_ratio": ot / max(ct, 1), "fp16_fraction": tot_fp16 / max(ct, 1),
                        "rtvq_fraction": tot_rtvq / max(ct, 1), "bases_fraction": tot_bases / max(ct, 1),
                        "num_parameters": len(stats["per_parameter"]), "num_tasks": num_tasks, "num_bits": num_bits,
                        "num_stages": num_stages}
    return stats


def compute_compression_statistics(task_vectors, compressed_all, bases, config) -> Dict:
    """diagnostics.py:385-566 on the reference's structures."""
    names = list(task_vectors.keys())
    first = task_vectors[names[0]]
    numel = {p: int(v.numel()) for p, v in first.items()}
    n_with = {p: sum(1 for tv in task_vectors.values() if p in tv) for p in first}
    sizes = {}
    for p, basis in bases.items():
        if p not in compressed_all:
            continue
        bm = basis.get("masked")
        if bm is None:
            sizes[p] = {"k": 0, "r": 0, "D": 0, "n_tasks": 0}
            continue
        nt = sum(1 for a in compressed_all[p].values() if a is not None and a.get("masked") is not None)
        sizes[p] = {"k": bm["k"], "r": bm["U_high"].shape[1] + bm["U_low"].shape[1], "D": bm["D"], "n_tasks": nt}
    stats = compression_statistics_from_sizes(numel, n_with, sizes, len(task_vectors), config.svd_low_bits,
                                              config.svd_rtvq_stages, names)
    for t, tv in task_vectors.items():
        stats["original"]["per_task_bytes"][t] = sum(d.numel() * 4 for d in tv.values())
    stats["original"]["total_bytes"] = sum(stats["original"]["per_task_bytes"].values())
    return stats


def print_detailed_compression_report(compression_stats: Dict, config, top_n_params: int = 5) -> None:
    s = compression_stats.get("summary", {})
    print(f"   compression: {s.get('original_size_mb', 0):.2f} MB -> {s.get('compressed_size_mb', 0):.2f} MB "
          f"(x{s.get('overall_compression_ratio', 0):.2f}); fp16 {s.get('fp16_fraction', 0):.1%}, "
          f"RTVQ {s.get('rtvq_fraction', 0):.1%}, bases {s.get('bases_fraction', 0):.1%}")
    top = sorted(compression_stats.get("per_parameter", {}).items(), key=lambda kv: kv[1].get("original_bytes", 0),
                 reverse=True)[:top_n_params]
    for name, ps in top:
        print(f"     {name[:50]:50s} k={ps.get('k', 0):3d} D={ps.get('D', 0):9d} x{ps.get('compression_ratio', 0):.2f}")


def print_diagnostics_summary(diagnostics: Dict) -> None:
    s = diagnostics.get("summary", {})
    print(f"   diagnostics: {s.get('num_parameters', 0)} parameters, average rank {s.get('average_rank', 0):.2f} "
          f"(+-{s.get('std_rank', 0):.2f}), energy {s.get('average_energy_retained', 0):.4f}, "
          f"reconstruction error {s.get('average_reconstruction_error', 0):.6f}, "
          f"compression ratio {s.get('average_compression_ratio', 0):.2f}")


def compute_coefficient_histograms(compressed_params: Dict[str, Dict], quantizer: RTVQQuantizer, num_bins: int = 50,
                                   device: str = "cpu") -> Dict:
    """diagnostics.py:324-382 (not used by the pipeline)."""
    hi, lo = [], []
    for art in compressed_params.values():
        if art.get("masked") is None:
            continue
        hi.append(art["masked"]["c_high_fp16"].float().flatten().cpu())
        lo.append(quantizer.dequantize(art["masked"]["c_low_quant"], device="cpu").float().flatten())
    if not hi:
        return {}

    def hist(parts):
        a = np.abs(torch.cat(parts).numpy())
        counts, edges = np.histogram(a, bins=num_bins)
        return {"counts": counts.tolist(), "bin_edges": edges.tolist(), "mean": float(a.mean()), "std": float(a.std()),
                "max": float(a.max())}
    return {"c_high": hist(hi), "c_low": hist(lo)}
