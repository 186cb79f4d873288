"""Task weighting (host logic; N scalars).  Mirrors src/svd_hybrid/weighting.py:61-396."""
import json
from typing import Dict, List, Optional

import numpy as np
import torch


def load_performance_metrics(performance_file: str, task_names: List[str]) -> Dict[str, float]:
    """weighting.py:61-117: exact key, else case/underscore/hyphen-insensitive key, else 1.0."""
    with open(performance_file, "r") as f:
        table = json.load(f)

    def norm(s: str) -> str:
        return s.lower().replace("_", "").replace("-", "")

    out = {}
    for name in task_names:
        if name in table:
            out[name] = float(table[name])
            continue
        hit = next((v for k, v in table.items() if norm(k) == norm(name)), None)
        if hit is None:
            print(f"Warning: No performance metric found for {name}, using 1.0")
            hit = 1.0
        out[name] = float(hit)
    return out


def compute_uniform_weights(task_names: List[str]) -> Dict[str, float]:
    w = 1.0 / len(task_names)
    return {n: w for n in task_names}


def compute_performance_weights(performance_metrics: Dict[str, float], temperature: float = 1.0) -> Dict[str, float]:
    """softmax(acc / T) evaluated in fp32 like the reference (weighting.py:182-191)."""
    if not performance_metrics:
        return {}
    names = list(performance_metrics.keys())
    perf = torch.tensor([performance_metrics[n] for n in names])
    w = torch.softmax(perf / temperature, dim=0)
    return {n: x.item() for n, x in zip(names, w)}


def compute_cluster_weights(task_names: List[str], cluster_assignments: Dict[str, int],
                            cluster_performance: Optional[Dict[int, float]] = None) -> Dict[str, float]:
    """Equal (or softmax-of-performance) weight per cluster, split equally inside (weighting.py:232-263)."""
    counts: Dict[int, int] = {}
    for n in task_names:
        c = cluster_assignments.get(n, 0)
        counts[c] = counts.get(c, 0) + 1
    if cluster_performance is not None:
        ids = list(counts.keys())
        sm = torch.softmax(torch.tensor([cluster_performance.get(c, 1.0) for c in ids]), dim=0)
        cw = {c: x.item() for c, x in zip(ids, sm)}
    else:
        cw = {c: 1.0 / len(counts) for c in counts}
    out = {n: cw[cluster_assignments.get(n, 0)] / counts[cluster_assignments.get(n, 0)] for n in task_names}
    tot = sum(out.values())
    if tot > 0:
        out = {k: v / tot for k, v in out.items()}
    return out


def compute_weights(task_names: List[str], weighting_strategy: str = "uniform", performance_file: Optional[str] = None,
                    temperature: float = 1.0, cluster_assignments: Optional[Dict[str, int]] = None,
                    cluster_performance: Optional[Dict[int, float]] = None) -> Dict[str, float]:
    if weighting_strategy == "uniform":
        return compute_uniform_weights(task_names)
    if weighting_strategy == "performance":
        if performance_file is None:
            print("Warning: Performance weighting requested but no performance file provided")
            return compute_uniform_weights(task_names)
        return compute_performance_weights(load_performance_metrics(performance_file, task_names), temperature)
    if weighting_strategy == "cluster":
        if cluster_assignments is None:
            print("Warning: Cluster weighting requested but no cluster assignments provided")
            return compute_uniform_weights(task_names)
        return compute_cluster_weights(task_names, cluster_assignments, cluster_performance)
    raise ValueError(f"Unknown weighting strategy: {weighting_strategy}")


def apply_weights_to_tensors(tensors: Dict, weights: Dict, device: str = "cpu") -> torch.Tensor:
    """Weighted average of a few same-shape tensors in sorted-key order (weighting.py:332-372).
    Plumbing for the fine-grained API (cluster results); the fused path folds this into K2."""
    if not tensors:
        raise ValueError("Empty tensor dictionary")
    keys = sorted(tensors.keys())
    st = torch.stack([tensors[k].to(device).float() for k in keys], dim=0)
    w = torch.tensor([weights.get(k, 1.0 / len(keys)) for k in keys], device=device, dtype=torch.float32)
    w = (w / w.sum()).view([len(keys)] + [1] * (st.ndim - 1))
    return (st * w).sum(dim=0)


def get_weight_statistics(weights: Dict[str, float]) -> Dict[str, float]:
    if not weights:
        return {}
    v = list(weights.values())
    return {"min": min(v), "max": max(v), "mean": sum(v) / len(v), "std": np.std(v),
            "entropy": -sum(w * np.log(w + 1e-10) for w in v)}


def cluster_omega(weights: Dict[str, float], cluster_assignments: Dict[str, int]) -> Dict[int, float]:
    """Cross-cluster weights of merge_with_clustering: every cluster is scored by the mean weight of its members
    (src/svd_hybrid/merge.py:614-618), the scores go through an fp32 softmax in cluster-dict order and the fp32
    renormalisation of apply_weights_to_tensors in sorted-id order (src/svd_hybrid/clustering.py:399-423,
    weighting.py:362-366)."""
    clusters: Dict[int, List[str]] = {}
    for n, c in cluster_assignments.items():
        clusters.setdefault(c, []).append(n)
    ids = list(clusters.keys())
    score = np.asarray([sum(weights.get(n, 1.0) for n in clusters[c]) / len(clusters[c]) for c in ids], np.float32)
    # numpy instead of torch.softmax keeps this off the critical path (a last-ulp difference in exp changes
    # a weight by ~1e-8 relative)
    ex = np.exp(score - score.max(), dtype=np.float32)
    sm = (ex / ex.sum(dtype=np.float32)).astype(np.float32)
    order = sorted(range(len(ids)), key=lambda i: ids[i])
    wt = sm[order]
    wt = (wt / wt.sum(dtype=np.float32)).astype(np.float32)
    return {ids[i]: float(wt[j]) for j, i in enumerate(order)}


def effective_merge_weights(task_names: List[str], weights: Dict[str, float],
                            cluster_assignments: Optional[Dict[str, int]] = None) -> Dict[str, float]:
    """Per-task weights that make ONE weighted average equal to the reference's merge WHEN EVERY TASK HAS THE
    PARAMETER (the per-parameter form, which also covers missing parameters, is K2's average_param fed by
    cluster_omega; this closed form is what tests compare it with).

    Without clustering this is ``weights``.  With clustering the reference merges per cluster with
    member weights renormalised inside the cluster, scores each cluster by the mean member weight,
    and averages the per-cluster results with softmax(score) (src/svd_hybrid/merge.py:586-626,
    src/svd_hybrid/clustering.py:399-423).  Every step is linear in the coefficients, so the result
    equals a single average with w_eff[t] = softmax(score)[c(t)] * w[t] / sum_{s in c(t)} w[s].
    """
    if not cluster_assignments:
        return {n: weights.get(n, 1.0 / len(task_names)) for n in task_names}
    clusters: Dict[int, List[str]] = {}
    for n, c in cluster_assignments.items():
        clusters.setdefault(c, []).append(n)
    omega = cluster_omega(weights, cluster_assignments)
    out = {}
    for c, members in clusters.items():
        tot = sum(weights.get(n, 1.0) for n in members)
        for n in members:
            out[n] = omega[c] * (weights.get(n, 1.0) / tot)
    return {n: out.get(n, 0.0) for n in task_names}
