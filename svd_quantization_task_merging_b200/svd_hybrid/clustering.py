"""Task clustering for cluster weighting.  Mirrors src/svd_hybrid/clustering.py:55-425.

The reference flattens every task vector into an [N x P_total] fp32 matrix on the host and runs
sklearn KMeans on it (742 s of an 863 s ViT-L-14 run).  K-means only sees pairwise geometry, so
the N x N whole-model task Gram -- a free by-product of the K1 streaming pass -- is enough: the
reference's k-means procedure (KMeans(k, random_state=42, n_init=10): numpy RandomState stream,
k-means++ seeding, float32 Lloyd iterations, best-of-10) is run on an N x N embedding E with
E E^T = normalised Gram by libsvdq's host routine svdq_host_kmeans (~30 us; label-identical to
the sklearn call on the same embedding, tests/test_kmeans_native.py).
"""
import functools
from typing import Dict, List, Tuple

import numpy as np
import torch


def embedding_from_gram(gram: np.ndarray) -> np.ndarray:
    """Rows L2-normalised exactly as clustering.py:232 (x / (||x|| + 1e-8)), embedded isometrically."""
    g = np.asarray(gram, np.float64)
    nrm = np.sqrt(np.clip(np.diag(g), 0.0, None)) + 1e-8
    gn = g / nrm[:, None] / nrm[None, :]
    lam, vec = np.linalg.eigh((gn + gn.T) * 0.5)
    return (vec * np.sqrt(np.clip(lam, 0.0, None))[None, :]).astype(np.float32)


def compute_kmeans_clustering(features: np.ndarray, k: int, random_state: int = 42) -> np.ndarray:
    """clustering.py:123-156: KMeans(n_clusters=k, random_state=random_state, n_init=10).fit_predict(features),
    restated in libsvdq (svdq_host_kmeans): same labels, without sklearn's per-call overhead."""
    features = np.asarray(features)
    if k <= 0 or k > features.shape[0]:
        raise ValueError(f"Invalid k={k} for {features.shape[0]} samples")
    from .. import _native
    f = np.ascontiguousarray(features, np.float32)
    labels = np.empty(f.shape[0], np.int32)
    _native.call("svdq_host_kmeans", f.ctypes.data, f.shape[0], f.shape[1], int(k), int(random_state) & 0xFFFFFFFF,
                 10, 300, 1e-4, labels.ctypes.data, None)
    return labels


def compute_kmeans_clustering_sklearn(features: np.ndarray, k: int, random_state: int = 42) -> np.ndarray:
    """The reference's very call (cross-check for compute_kmeans_clustering; needs scikit-learn)."""
    if k <= 0 or k > features.shape[0]:
        raise ValueError(f"Invalid k={k} for {features.shape[0]} samples")
    from sklearn.cluster import KMeans
    return KMeans(n_clusters=k, random_state=random_state, n_init=10).fit_predict(features)


def compute_hierarchical_clustering(features: np.ndarray, k: int, method: str = "ward") -> np.ndarray:
    if k <= 0 or k > features.shape[0]:
        raise ValueError(f"Invalid k={k} for {features.shape[0]} samples")
    from scipy.cluster.hierarchy import fcluster, linkage
    return fcluster(linkage(features, method=method), k, criterion="maxclust") - 1


def normalised_gram(gram: np.ndarray) -> np.ndarray:
    g = np.asarray(gram, np.float64)
    nrm = np.sqrt(np.clip(np.diag(g), 0.0, None)) + 1e-8
    return g / nrm[:, None] / nrm[None, :]


def _inertia(labels: np.ndarray, g: np.ndarray, k: int) -> np.ndarray:
    """k-means objective of M labelings [M x N] computed from the Gram matrix alone:
    sum_c ( sum_{i in c} G_ii - (1/|c|) sum_{i,j in c} G_ij )."""
    diag = np.diag(g)
    tot = np.zeros(labels.shape[0])
    for c in range(k):
        ind = (labels == c).astype(np.float64)
        cnt = ind.sum(1)
        quad = ((ind @ g) * ind).sum(1)
        tot += ind @ diag - quad / np.maximum(cnt, 1.0)
    return tot


@functools.lru_cache(maxsize=16)
def _restricted_growth(n: int, k: int, limit: int):
    """All set partitions of n items into <= k blocks as label rows (first item in block 0);
    None when there are more than ``limit`` of them."""
    rows = np.zeros((1, 1), np.int8)
    for _ in range(1, n):
        mx = rows.max(1)
        reps = np.minimum(mx + 2, k)                      # labels 0 .. min(max+1, k-1)
        if int(reps.sum()) > limit:
            return None
        idx = np.repeat(np.arange(rows.shape[0]), reps)
        start = np.cumsum(reps) - reps
        new = (np.arange(idx.size) - np.repeat(start, reps)).astype(np.int8)
        rows = np.concatenate([rows[idx], new[:, None]], 1)
    return rows


@functools.lru_cache(maxsize=16)
def _partition_table(n: int, k: int, limit: int):
    """Cached enumeration for the exact solver: candidate labelings, their one-hot indicator matrices
    [k x M x n] and cluster sizes [k x M] (None when the enumeration would exceed ``limit``)."""
    rows = _restricted_growth(n, k, limit)
    if rows is None:
        return None
    used = rows.max(1) + 1
    cand = rows[used == min(k, n)] if (used == min(k, n)).any() else rows
    ind = np.stack([(cand == c).astype(np.float64) for c in range(k)], 0)
    cnt = np.maximum(ind.sum(2), 1.0)
    return cand.astype(np.int64), ind, cnt


def kmeans_partition_from_gram(gram_normalised: np.ndarray, k: int, exact_limit: int = 200_000,
                               n_init: int = 64, seed: int = 42) -> np.ndarray:
    """Deterministic k-means on points known only through their Gram matrix.

    Small problems (the usual 8-20 tasks, k = 2) are solved EXACTLY by enumerating all set partitions;
    larger ones by Lloyd iterations from ``n_init`` deterministic seedings, keeping the lowest inertia.
    Labels are canonical: clusters are numbered by first appearance.
    """
    g = np.asarray(gram_normalised, np.float64)
    n = g.shape[0]
    if k <= 0 or k > n:
        raise ValueError(f"Invalid k={k} for {n} samples")
    table = _partition_table(n, k, exact_limit)
    if table is not None:
        cand, ind, cnt = table
        # inertia = sum_c ( sum_{i in c} G_ii - (1/|c|) sum_{i,j in c} G_ij ); the first term is trace(G) for all
        quad = (np.matmul(ind, g) * ind).sum(2)
        return cand[int(np.argmax((quad / cnt).sum(0)))]
    rng = np.random.default_rng(seed)
    diag = np.diag(g)
    best, best_val = None, np.inf
    for _ in range(n_init):
        lab = rng.integers(0, k, n)
        lab[rng.permutation(n)[:k]] = np.arange(k)
        for _ in range(100):
            ind = (lab[None, :] == np.arange(k)[:, None]).astype(np.float64)          # [k x n]
            cnt = np.maximum(ind.sum(1), 1.0)
            cross = ind @ g                                                          # [k x n]
            cc = (cross * ind).sum(1)
            dist = diag[None, :] - 2.0 * cross / cnt[:, None] + (cc / cnt ** 2)[:, None]
            new = dist.argmin(0)
            if np.array_equal(new, lab):
                break
            lab = new
        val = float(_inertia(lab[None, :], g, k)[0])
        if val < best_val - 1e-15:
            best, best_val = lab.copy(), val
    remap, out = {}, np.zeros(n, np.int64)
    for i, c in enumerate(best):
        out[i] = remap.setdefault(int(c), len(remap))
    return out


def cluster_from_gram(gram: np.ndarray, task_names_in_gram_order: List[str], k: int,
                      method: str = "kmeans", backend: str = "kmeans") -> Dict[str, int]:
    """Cluster tasks from the whole-model task Gram.  Rows are re-ordered to sorted task names
    first, because the reference builds its feature matrix in sorted order (clustering.py:87).

    backend "kmeans" (default): the reference's procedure, KMeans(k, random_state=42, n_init=10), restated in
                       libsvdq (svdq_host_kmeans) on an isometric N x N embedding -- the labels sklearn gives;
    backend "sklearn": the same call made through scikit-learn itself (8-23 ms; cross-check);
    backend "exact"  : opt-in, NOT the reference's algorithm: the global k-means optimum by enumeration.  It
                       differs from the reference partition whenever sklearn's best-of-10 stops in a local
                       optimum (about one unclustered 8-task input in five).
    """
    names = list(task_names_in_gram_order)
    order = sorted(range(len(names)), key=lambda i: names[i])
    g = np.asarray(gram, np.float64)[np.ix_(order, order)]
    if method == "kmeans" and backend == "exact":
        labels = kmeans_partition_from_gram(normalised_gram(g), k)
        return {names[i]: int(l) for i, l in zip(order, labels)}
    feats = embedding_from_gram(g)
    if method == "kmeans" and backend == "sklearn":
        labels = compute_kmeans_clustering_sklearn(feats, k)
    elif method == "kmeans":
        if backend != "kmeans":
            raise ValueError(f"Unknown cluster backend: {backend}")
        labels = compute_kmeans_clustering(feats, k)
    elif method == "hierarchical":
        labels = compute_hierarchical_clustering(feats, k)
    else:
        raise ValueError(f"Unknown clustering method: {method}")
    return {names[i]: int(l) for i, l in zip(order, labels)}


def cluster_tasks(task_vectors: Dict[str, Dict[str, torch.Tensor]], k: int, method: str = "kmeans") -> Dict[str, int]:
    """Reference signature (clustering.py:198).  The Gram is accumulated on the GPU by the K1 kernel."""
    from ..engine import task_vector_gram
    names = list(task_vectors.keys())
    return cluster_from_gram(task_vector_gram(task_vectors, names), names, k, method)


def get_cluster_members(cluster_assignments: Dict[str, int]) -> Dict[int, List[str]]:
    out: Dict[int, List[str]] = {}
    for name, cid in cluster_assignments.items():
        out.setdefault(cid, []).append(name)
    return out


def flatten_task_vectors(task_vectors: Dict[str, Dict[str, torch.Tensor]]) -> Tuple[np.ndarray, List[str]]:
    """[N x P_total] host feature matrix in sorted task / sorted parameter order, zero-filled where a task lacks
    a parameter (clustering.py:55-120).  Kept for API parity; the pipeline clusters from the Gram instead."""
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
        rows.append(torch.cat(parts, dim=0).cpu().numpy())
    return np.stack(rows, axis=0), names


def compute_cluster_statistics(task_vectors: Dict[str, Dict[str, torch.Tensor]],
                               cluster_assignments: Dict[str, int]) -> Dict[int, Dict]:
    """Per-cluster size / members / distances to the centroid (clustering.py:278-316)."""
    feats, names = flatten_task_vectors(task_vectors)
    idx = {n: i for i, n in enumerate(names)}
    out = {}
    for cid, members in get_cluster_members(cluster_assignments).items():
        f = feats[[idx[m] for m in members]]
        d = np.linalg.norm(f - f.mean(axis=0), axis=1)
        out[cid] = {"size": len(members), "members": members, "mean_distance_to_centroid": float(d.mean()),
                    "max_distance_to_centroid": float(d.max()), "min_distance_to_centroid": float(d.min())}
    return out


def merge_by_cluster(task_vectors: Dict[str, Dict[str, torch.Tensor]], cluster_assignments: Dict[str, int],
                     weights: Dict[str, float], device: str = "cpu") -> Dict[int, Dict[str, torch.Tensor]]:
    """Weighted average of raw task vectors inside each cluster (clustering.py:319-371)."""
    from .weighting import apply_weights_to_tensors
    out = {}
    for cid, members in get_cluster_members(cluster_assignments).items():
        cw = {n: weights.get(n, 1.0) for n in members}
        tot = sum(cw.values())
        cw = {n: v / tot for n, v in cw.items()}
        params = set()
        for n in members:
            params.update(task_vectors[n].keys())
        out[cid] = {p: apply_weights_to_tensors({n: task_vectors[n][p] for n in members if p in task_vectors[n]}, cw, device)
                    for p in params}
    return out


def merge_cluster_results(cluster_merged: Dict[int, Dict[str, torch.Tensor]], cluster_performance: Dict[int, float],
                          device: str = "cpu") -> Dict[str, torch.Tensor]:
    """softmax(cluster score)-weighted average of per-cluster results (clustering.py:374-425)."""
    from .weighting import apply_weights_to_tensors
    if not cluster_merged:
        return {}
    ids = list(cluster_merged.keys())
    if cluster_performance:
        w = torch.softmax(torch.tensor([cluster_performance.get(c, 1.0) for c in ids]), dim=0)
    else:
        w = torch.ones(len(ids)) / len(ids)
    wd = {c: x.item() for c, x in zip(ids, w)}
    params = set()
    for d in cluster_merged.values():
        params.update(d.keys())
    return {p: apply_weights_to_tensors({c: cluster_merged[c][p] for c in ids if p in cluster_merged[c]}, wd, device)
            for p in params}
