"""Torch-CPU restatement of the reference SVD-Hybrid merge path (TEST ORACLE).

See ``oracle/__init__.py`` for who may import this.  Every function names the
reference lines it restates (paths relative to /root/reference).  The code is
written from the algorithm, operating on plain dicts of CPU tensors, and keeps
the reference's operation ORDER wherever order changes bits (separate fp32
multiply and add in the quantiser, half-to-even rounding, double-accumulated
cumsum, fp16 cast of the basis before projection, left-to-right stage sums).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

Tensor = torch.Tensor


# --------------------------------------------------------------------------
# configuration (mirror of the fields the hot path reads)
# --------------------------------------------------------------------------
@dataclass
class RefConfig:
    """Subset of SVDHybridConfig (src/svd_hybrid/config.py:159-205) used by the path."""
    tasks: List[str] = field(default_factory=list)
    svd_energy_threshold: float = 0.95
    svd_max_rank: Optional[int] = 64
    svd_center: bool = True
    svd_fp16: bool = True
    svd_low_bits: int = 4
    svd_rtvq_stages: int = 2
    svd_mask_strategy: str = "union"
    svd_include_noise: bool = False
    svd_noise_shrink: float = 0.5
    svd_weighting: str = "uniform"
    svd_weighting_temperature: float = 5.0
    svd_cluster_k: int = 2
    svd_min_mask_size: int = 10
    svd_eval_reconstruction: bool = True
    performance: Optional[Dict[str, float]] = None   # contents of performance_file


# --------------------------------------------------------------------------
# a1  task vectors            src/svd_hybrid/task_vector_loader.py:103-145
# --------------------------------------------------------------------------
def task_vector(base: Dict[str, Tensor], finetuned: Dict[str, Tensor]) -> Dict[str, Tensor]:
    out = {}
    for name, b in base.items():
        f = finetuned.get(name)
        if f is None or f.shape != b.shape:
            continue
        out[name] = (f - b).detach()
    return out


# --------------------------------------------------------------------------
# a2  tall-mask combination   src/svd_hybrid/mask_loader.py:412-485,488-648
# --------------------------------------------------------------------------
def combine_mask_list(masks: Sequence[Tensor], strategy: str) -> Tensor:
    if len(masks) == 0:
        raise ValueError("Empty mask list")
    if strategy == "union":
        acc = masks[0].clone()
        for m in masks[1:]:
            acc = acc | m
        return acc
    if strategy == "intersection":
        acc = masks[0].clone()
        for m in masks[1:]:
            acc = acc & m
        return acc
    if strategy == "majority":
        votes = torch.stack([m.float() for m in masks], 0).sum(0)
        return votes >= (0.5 * len(masks))
    raise ValueError(f"Unknown mask strategy: {strategy}")


def combine_masks(task_masks: Dict[str, Optional[Dict[str, Tensor]]], strategy: str) -> Dict[str, Tensor]:
    if strategy not in ("union", "intersection", "majority"):
        raise ValueError(f"Unknown mask strategy: {strategy}")
    names = set()
    for pm in task_masks.values():
        if pm is not None:
            names.update(pm.keys())
    out = {}
    for name in names:
        present = [pm[name] for pm in task_masks.values() if pm is not None and name in pm]
        if present:
            out[name] = combine_mask_list(present, strategy)
    return out


# --------------------------------------------------------------------------
# a7  energy spectrum / rank  src/svd_hybrid/basis.py:116-213
# --------------------------------------------------------------------------
def energy_spectrum(S: Tensor) -> Tensor:
    e = S ** 2
    tot = e.sum()
    if tot < 1e-10:
        return torch.ones_like(e)
    return torch.cumsum(e, 0) / tot


def select_rank(S: Tensor, thr: float = 0.90, max_rank: Optional[int] = None, min_rank: int = 1) -> int:
    k = int((energy_spectrum(S) < thr).sum().item()) + 1
    k = max(k, min_rank)
    if max_rank is not None:
        k = min(k, max_rank)
    return min(k, len(S))


# --------------------------------------------------------------------------
# a5-a8  basis                src/svd_hybrid/basis.py:63-113,216-249,252-409
# --------------------------------------------------------------------------
def build_basis(deltas: Sequence[Tensor], thr: float, max_rank: Optional[int], center: bool) -> Dict:
    if len(deltas) == 0:
        raise ValueError("Empty delta list")
    T = torch.stack(list(deltas), dim=1)
    mean = None
    if center:
        mean = T.mean(dim=1, keepdim=True)
        T = T - mean
    D, N = T.shape
    U, S, Vh = torch.linalg.svd(T, full_matrices=False)
    k = select_rank(S, thr, max_rank)
    return {
        "U_high": U[:, :k].contiguous(),
        "U_low": U[:, k:].contiguous(),
        "singular_values": S,
        "k": k,
        "mean": mean,
        "energy_retained": energy_spectrum(S)[k - 1].item() if k > 0 else 0,
        "D": D,
        "N": N,
        "Vh": Vh,          # oracle-only extra: lets the harness align signs
    }


# --------------------------------------------------------------------------
# a11-a13  RTVQ               src/svd_hybrid/rtvq.py:4-139 (= quantization_utils.py:76-172)
# --------------------------------------------------------------------------
def asym_quant(X: Tensor, bits: int = 8) -> Tuple[Tensor, Tensor, Tensor]:
    lo, hi = X.min(), X.max()
    qmax = 2 ** bits - 1
    scale = qmax / (hi - lo)
    zp = -1 * torch.round(scale * lo)
    q = torch.round(scale * X + zp).clamp(0, qmax)
    return q.to(torch.uint8 if bits <= 8 else torch.int16), scale, zp


def asym_dequant(q: Tensor, scale: Tensor, zp: Tensor) -> Tensor:
    return (q.float() - zp) / scale


def rtvq_quantize(x: Tensor, bits: int = 4, stages: int = 2) -> List[Dict]:
    if x.numel() == 0:
        return []
    res = x.clone()
    out = []
    for s in range(stages):
        norm_before = res.norm().item()
        q, scale, zp = asym_quant(res, bits)
        res = res - asym_dequant(q, scale, zp)
        out.append({"stage": s, "quantized": q, "scale": scale, "zero_point": zp,
                    "residual_norm": norm_before})
    return out


def rtvq_dequantize(payloads: List[Dict]) -> Tensor:
    if not payloads:
        return torch.tensor([])
    acc = torch.zeros_like(payloads[0]["quantized"].float())
    for p in payloads:
        acc = acc + asym_dequant(p["quantized"], p["scale"], p["zero_point"])
    return acc


def rtvq_pack(x: Tensor, bits: int, stages: int) -> Dict:
    """RTVQQuantizer.quantize (rtvq.py:111-126)."""
    return {"payloads": rtvq_quantize(x, bits, stages), "num_bits": bits, "num_stages": stages,
            "original_shape": x.shape, "original_dtype": str(x.dtype)}


def rtvq_unpack(obj: Dict) -> Tensor:
    """RTVQQuantizer.dequantize (rtvq.py:128-139)."""
    out = rtvq_dequantize(obj["payloads"])
    if "original_shape" in obj:
        out = out.view(obj["original_shape"])
    return out


def compression_ratio(n: int, bits: int, stages: int) -> float:
    """estimate_compression_ratio (rtvq.py:142-161) as a function of the element count."""
    return (n * 4) / max(n * bits / 8 * stages + 8 * stages, 1)


# root quantization_utils.py:60-73,102-134
def absmax_quant(X: Tensor, bits: int = 8) -> Tuple[Tensor, Tensor]:
    s = (2 ** (bits - 1) - 1) / torch.max(torch.abs(X))
    return (s * X).round().to(torch.int8 if bits <= 8 else torch.int16), s


def absmax_dequant(q: Tensor, s: Tensor) -> Tensor:
    return q.float() * s      # (sic) the reference multiplies


# --------------------------------------------------------------------------
# a9-a10  projection          src/svd_hybrid/compress.py:6-56, cli.py:355-361
# --------------------------------------------------------------------------
def compress_task(delta: Tensor, U_high: Tensor, U_low: Tensor, mean: Optional[Tensor],
                  bits: int, stages: int) -> Dict:
    x = delta if mean is None else delta - mean.squeeze()
    x = x.float()
    c_high = U_high.float().T @ x
    c_low = U_low.float().T @ x
    return {"c_high_fp16": c_high.half(), "c_low_quant": rtvq_pack(c_low, bits, stages),
            "c_low_fp32": c_low, "c_high_fp32": c_high}    # *_fp32: oracle-only extras


# --------------------------------------------------------------------------
# a14  weights                src/svd_hybrid/weighting.py:120-329
# --------------------------------------------------------------------------
def task_weights(cfg: RefConfig, assignments: Optional[Dict[str, int]] = None) -> Dict[str, float]:
    names = list(cfg.tasks)
    if cfg.svd_weighting == "performance" and cfg.performance is not None:
        perf = torch.tensor([float(cfg.performance.get(n, 1.0)) for n in names])
        w = torch.softmax(perf / cfg.svd_weighting_temperature, 0)
        return {n: x.item() for n, x in zip(names, w)}
    if cfg.svd_weighting == "cluster" and assignments is not None:
        counts: Dict[int, int] = {}
        for n in names:
            c = assignments.get(n, 0)
            counts[c] = counts.get(c, 0) + 1
        per_cluster = 1.0 / len(counts)
        w = {n: per_cluster / counts[assignments.get(n, 0)] for n in names}
        tot = sum(w.values())
        return {n: v / tot for n, v in w.items()} if tot > 0 else w
    return {n: 1.0 / len(names) for n in names}


# --------------------------------------------------------------------------
# a15  clustering             src/svd_hybrid/clustering.py:55-156,198-245
# --------------------------------------------------------------------------
def cluster_tasks_full(task_vectors: Dict[str, Dict[str, Tensor]], k: int) -> Dict[str, int]:
    """Full-feature k-means exactly as the reference runs it (slow, memory hungry)."""
    from sklearn.cluster import KMeans
    names = sorted(task_vectors.keys())
    params = sorted({p for tv in task_vectors.values() for p in tv})
    rows = []
    for n in names:
        tv = task_vectors[n]
        parts = []
        for p in params:
            if p in tv:
                parts.append(tv[p].flatten())
            else:
                ref = next(t[p] for t in task_vectors.values() if p in t)
                parts.append(torch.zeros_like(ref).flatten())
        rows.append(torch.cat(parts).cpu().numpy())
    F = np.stack(rows, 0)
    F = F / (np.linalg.norm(F, axis=1, keepdims=True) + 1e-8)
    if k <= 0 or k > F.shape[0]:
        raise ValueError(f"Invalid k={k} for {F.shape[0]} samples")
    labels = KMeans(n_clusters=k, random_state=42, n_init=10).fit_predict(F)
    return {n: int(l) for n, l in zip(names, labels)}


# --------------------------------------------------------------------------
# a16-a19  merge              src/svd_hybrid/merge.py:61-426,555-626; clustering.py:374-425
# --------------------------------------------------------------------------
def average_coeffs(per_task: Dict[str, Dict], weights: Dict[str, float]) -> Tuple[Optional[Tensor], Optional[Tensor]]:
    names = sorted(per_task.keys())
    hi, lo, w = [], [], []
    for n in names:
        art = per_task[n]
        if art is None:
            continue
        hi.append(art["c_high_fp16"].float())
        lo.append(rtvq_unpack(art["c_low_quant"]).float())
        w.append(weights.get(n, 1.0 / len(names)))
    if not hi:
        return None, None
    tot = sum(w)
    wt = torch.tensor([x / tot for x in w], dtype=torch.float32).view(-1, 1)
    return (torch.stack(hi, 0) * wt).sum(0), (torch.stack(lo, 0) * wt).sum(0)


def reconstruct(c_hi: Tensor, c_lo: Tensor, basis: Dict) -> Tensor:
    out = basis["U_high"].float() @ c_hi + basis["U_low"].float() @ c_lo
    if basis["mean"] is not None:
        out = out + basis["mean"].squeeze().float()
    return out


def scatter_masked(vals: Tensor, mask: Tensor, shape, unmasked_vals: Optional[Tensor] = None) -> Tensor:
    """reconstruct_from_masked (mask_loader.py:712-763); unmasked_vals = the noise region (:757-760)."""
    flat = mask.flatten()
    out = torch.zeros_like(flat, dtype=vals.dtype)
    out[flat] = vals
    if unmasked_vals is not None:
        out[~flat] = unmasked_vals
    return out.view(shape)


def merge_deltas(compressed: Dict[str, Dict[str, Dict]], bases: Dict[str, Dict], masks: Dict[str, Tensor],
                 weights: Dict[str, float], shapes: Dict[str, torch.Size],
                 noise: Optional[Tuple[Dict, Dict, float]] = None) -> Dict[str, Tensor]:
    """merge_all_parameters / merge_parameter (merge.py:197-426).  ``noise`` = (compressed_noise, bases_noise,
    noise_shrink) when svd_include_noise: the unmasked positions get shrink * (U_n c_n + mean_n)
    (merge.py:257-284); without it they stay zero."""
    out = {}
    for name in sorted(compressed.keys()):
        c_hi, c_lo = average_coeffs(compressed[name], weights)
        if c_hi is None:
            out[name] = torch.zeros(shapes[name])
            continue
        vec = reconstruct(c_hi, c_lo, bases[name])
        rest = None
        if noise is not None and noise[1].get(name) is not None:
            n_hi, n_lo = average_coeffs(noise[0].get(name, {}), weights)
            if n_hi is not None:
                rest = reconstruct(n_hi, n_lo, noise[1][name]) * noise[2]
        m = masks.get(name)
        out[name] = scatter_masked(vec, m, shapes[name], rest) if m is not None else vec.view(shapes[name])
    return out


def weighted_stack(tensors: Dict, weights: Dict) -> Tensor:
    """apply_weights_to_tensors (weighting.py:332-372)."""
    keys = sorted(tensors.keys())
    st = torch.stack([tensors[k].float() for k in keys], 0)
    w = torch.tensor([weights.get(k, 1.0 / len(keys)) for k in keys], dtype=torch.float32)
    w = (w / w.sum()).view([len(keys)] + [1] * (st.ndim - 1))
    return (st * w).sum(0)


def merge_deltas_clustered(compressed, bases, masks, weights, assignments, shapes, noise=None) -> Dict[str, Tensor]:
    clusters: Dict[int, List[str]] = {}
    for n, c in assignments.items():
        clusters.setdefault(c, []).append(n)
    per_cluster, score = {}, {}
    for cid, members in clusters.items():
        cw = {n: weights.get(n, 1.0) for n in members}
        tot = sum(cw.values())
        cw = {n: v / tot for n, v in cw.items()}
        sub = {p: {n: art[n] for n in members if n in art} for p, art in compressed.items()}
        sub_noise = None
        if noise is not None:
            sub_noise = ({p: {n: art[n] for n in members if n in art} for p, art in noise[0].items()},
                         noise[1], noise[2])
        per_cluster[cid] = merge_deltas(sub, bases, masks, cw, shapes, sub_noise)
        score[cid] = sum(weights.get(n, 1.0) for n in members) / len(members)
    ids = list(per_cluster.keys())
    sm = torch.softmax(torch.tensor([score.get(c, 1.0) for c in ids]), 0)
    cwd = {c: w.item() for c, w in zip(ids, sm)}
    params = set()
    for d in per_cluster.values():
        params.update(d.keys())
    return {p: weighted_stack({c: per_cluster[c][p] for c in ids if p in per_cluster[c]}, cwd) for p in params}


# --------------------------------------------------------------------------
# a21  diagnostics            src/svd_hybrid/diagnostics.py:72-321
# --------------------------------------------------------------------------
def recon_error(orig: Tensor, rec: Tensor) -> Dict[str, float]:
    e = orig - rec
    on, en = orig.norm().item(), e.norm().item()
    return {"absolute_error": en, "relative_error": en / on if on > 1e-10 else 0,
            "max_absolute_error": e.abs().max().item(), "mean_absolute_error": e.abs().mean().item(),
            "original_norm": on, "reconstructed_norm": rec.norm().item()}


def diagnostics(task_vectors, compressed, bases, masks, cfg: RefConfig) -> Dict:
    out = {"config": {"svd_energy_threshold": cfg.svd_energy_threshold, "svd_max_rank": cfg.svd_max_rank,
                      "svd_low_bits": cfg.svd_low_bits, "svd_rtvq_stages": cfg.svd_rtvq_stages,
                      "svd_mask_strategy": cfg.svd_mask_strategy, "svd_weighting": cfg.svd_weighting},
           "per_parameter": {}, "summary": {}}
    first = next(iter(task_vectors.keys()))
    for name in sorted(bases.keys()):
        if name not in compressed:
            continue
        b = bases[name]
        d = {"param_name": name, "original_shape": None, "masked_size": 0, "unmasked_size": 0,
             "reconstruction_errors": {}, "compression_ratios": {}}
        out["per_parameter"][name] = d
        if b is None:
            continue
        if name in task_vectors[first]:
            d["original_shape"] = list(task_vectors[first][name].shape)
        m = masks.get(name)
        if m is not None:
            d["masked_size"] = int(m.sum().item())
            d["unmasked_size"] = int((~m).sum().item())
        else:
            d["masked_size"] = np.prod(d["original_shape"])
        d["basis"] = {"k": b["k"], "D": b["D"], "N": b["N"], "energy_retained": b["energy_retained"]}
        rel = []
        for t, tv in task_vectors.items():
            if name not in tv or t not in compressed[name] or compressed[name][t] is None:
                continue
            orig = tv[name]
            orig = orig.flatten()[m.flatten()] if (m is not None and m.shape == orig.shape) else orig.flatten()
            art = compressed[name][t]
            c_hi = art["c_high_fp16"].float()
            c_lo = rtvq_unpack(art["c_low_quant"]).float()
            rec = b["U_high"].float() @ c_hi + b["U_low"].float() @ c_lo      # (sic) no mean
            em = recon_error(orig, rec)
            rel.append(em["relative_error"])
            d["reconstruction_errors"][t] = em
            d["compression_ratios"][t] = compression_ratio(c_lo.numel(), art["c_low_quant"]["num_bits"],
                                                           art["c_low_quant"]["num_stages"])
        if rel:
            d["mean_relative_error"] = float(np.mean(rel))
            d["std_relative_error"] = float(np.std(rel))
            d["max_relative_error"] = float(np.max(rel))
            d["min_relative_error"] = float(np.min(rel))
    ranks, energy, errs, ratios = [], [], [], []
    for d in out["per_parameter"].values():
        if "basis" in d:
            ranks.append(d["basis"]["k"])
            energy.append(d["basis"]["energy_retained"])
        if "mean_relative_error" in d:
            errs.append(d["mean_relative_error"])
        if d.get("compression_ratios"):
            ratios.append(np.mean(list(d["compression_ratios"].values())))
    out["summary"] = {"num_parameters": len(out["per_parameter"]),
                      "average_rank": float(np.mean(ranks)) if ranks else 0,
                      "std_rank": float(np.std(ranks)) if ranks else 0,
                      "average_energy_retained": float(np.mean(energy)) if energy else 0,
                      "average_reconstruction_error": float(np.mean(errs)) if errs else 0,
                      "average_compression_ratio": float(np.mean(ratios)) if ratios else 0}
    return out


# --------------------------------------------------------------------------
# the whole path              src/svd_hybrid/cli.py:73-778 (steps 1-9, no disk)
# --------------------------------------------------------------------------
def run_reference_path(base: Dict[str, Tensor], finetuned: Dict[str, Dict[str, Tensor]],
                       task_masks: Optional[Dict[str, Optional[Dict[str, Tensor]]]], cfg: RefConfig,
                       assignments: Optional[Dict[str, int]] = None, stages_timing: Optional[Dict] = None) -> Dict:
    """Steps 1-9 of run_svd_hybrid_pipeline on in-memory state dicts.

    ``finetuned`` is keyed by task in ``cfg.tasks`` order.  ``assignments`` lets
    the caller inject a k-means partition (the full-feature clustering is
    minutes on ViT-L-14); when None and weighting == "cluster" the full
    reference clustering is run.  With cfg.svd_include_noise the rows outside
    the combined mask get their own basis / coefficients (result keys
    "bases_noise", "compressed_noise"; cli.py:336-351, basis.py:455-466,
    compress.py:91-107) and are merged with svd_noise_shrink (merge.py:257-284).
    """
    import time
    tick = time.perf_counter
    tm = stages_timing if stages_timing is not None else {}
    t0 = tick()
    tvs = {t: task_vector(base, finetuned[t]) for t in cfg.tasks}
    tm["task_vectors"] = tick() - t0

    t0 = tick()
    masks = combine_masks(task_masks, cfg.svd_mask_strategy) if task_masks else {}
    tm["masks"] = tick() - t0

    names = sorted({p for tv in tvs.values() for p in tv})
    shapes = {}
    for tv in tvs.values():
        for p, d in tv.items():
            shapes.setdefault(p, d.shape)

    # step 4: bases (cli.py:317-379)
    t0 = tick()
    bases: Dict[str, Optional[Dict]] = {}
    bases_noise: Dict[str, Optional[Dict]] = {}

    def half(b):
        if cfg.svd_fp16:
            b["U_high"] = b["U_high"].half()
            b["U_low"] = b["U_low"].half()
        return b

    for p in names:
        m = masks.get(p)
        cols, rest = [], []
        for t in cfg.tasks:
            if p not in tvs[t]:
                continue
            d = tvs[t][p]
            if m is not None and m.shape == d.shape:
                if m.sum() >= cfg.svd_min_mask_size:
                    cols.append(d.flatten()[m.flatten()])
                    if cfg.svd_include_noise:
                        rest.append(d.flatten()[~m.flatten()])       # get_unmasked_portion, mask_loader.py:682-709
            else:
                cols.append(d.flatten())
        if cols and len(cols[0]) > 0:
            bases[p] = half(build_basis(cols, cfg.svd_energy_threshold, cfg.svd_max_rank, cfg.svd_center))
            if cfg.svd_include_noise and rest and len(rest[0]) > 0:  # basis.py:455-466
                bases_noise[p] = half(build_basis(rest, cfg.svd_energy_threshold, cfg.svd_max_rank, cfg.svd_center))
    tm["basis"] = tick() - t0

    # step 5: compression (compress.py:114-207)
    t0 = tick()
    compressed: Dict[str, Dict[str, Dict]] = {}
    compressed_noise: Dict[str, Dict[str, Dict]] = {}
    for p in sorted(bases.keys()):
        b, m = bases[p], masks.get(p)
        bn = bases_noise.get(p)
        per_task, per_task_noise = {}, {}
        for t in cfg.tasks:
            if p not in tvs[t]:
                continue
            d = tvs[t][p]
            y = None
            if m is not None and m.shape == d.shape:
                x = d.flatten()[m.flatten()] if m.sum() >= cfg.svd_min_mask_size else torch.tensor([])
                if bn is not None:
                    y = d.flatten()[~m.flatten()]
            else:
                x = d.flatten()
            per_task[t] = (compress_task(x, b["U_high"], b["U_low"], b["mean"], cfg.svd_low_bits,
                                         cfg.svd_rtvq_stages) if len(x) > 0 else None)
            if y is not None and len(y) > 0:                         # compress.py:91-107
                per_task_noise[t] = compress_task(y, bn["U_high"], bn["U_low"], bn["mean"], cfg.svd_low_bits,
                                                  cfg.svd_rtvq_stages)
        if per_task:
            compressed[p] = per_task
        if per_task_noise:
            compressed_noise[p] = per_task_noise
    tm["compress"] = tick() - t0

    # step 6: weights
    t0 = tick()
    if cfg.svd_weighting == "cluster" and assignments is None:
        assignments = cluster_tasks_full(tvs, cfg.svd_cluster_k)
    weights = task_weights(cfg, assignments)
    tm["weights"] = tick() - t0

    # step 7-8: merge + apply
    t0 = tick()
    noise = (compressed_noise, bases_noise, cfg.svd_noise_shrink) if cfg.svd_include_noise else None
    if cfg.svd_weighting == "cluster" and assignments is not None:
        deltas = merge_deltas_clustered(compressed, bases, masks, weights, assignments, shapes, noise)
    else:
        deltas = merge_deltas(compressed, bases, masks, weights, shapes, noise)
    merged = {}
    for p, b in base.items():
        merged[p] = b + deltas[p] if p in deltas else b.clone()
    tm["merge"] = tick() - t0

    t0 = tick()
    diag = diagnostics(tvs, compressed, bases, masks, cfg) if cfg.svd_eval_reconstruction else {}
    diag["task_weights"] = weights
    if assignments:
        diag["cluster_assignments"] = assignments
    tm["diagnostics"] = tick() - t0
    return {"merged_state_dict": merged, "diagnostics": diag, "bases": bases, "compressed": compressed,
            "bases_noise": bases_noise, "compressed_noise": compressed_noise,
            "weights": weights, "cluster_assignments": assignments, "combined_masks": masks,
            "task_vectors": tvs, "merged_deltas": deltas}
