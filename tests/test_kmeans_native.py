"""svdq_host_kmeans (libsvdq, host side) against the call the reference makes,
sklearn.cluster.KMeans(n_clusters=k, random_state=42, n_init=10).fit_predict (reference:
src/svd_hybrid/clustering.py:123-156), and against the reference's full-feature cluster_tasks
(oracle.cluster_tasks_full, pinned to the reference golden).  CPU only: no CUDA call is made."""
import warnings

import numpy as np
import pytest
import torch

from oracle import svd_hybrid_ref as R
from svd_quantization_task_merging_b200.svd_hybrid import clustering


def _tasks(rng, kind, n, k, P=160):
    if kind == 0:        # unclustered
        return rng.standard_normal((n, P))
    if kind == 1:        # clustered
        c = rng.standard_normal((k, P))
        return c[rng.integers(0, k, n)] + 0.5 * rng.standard_normal((n, P))
    # a shared component plus task-specific parts of decaying size (what fine-tuned task vectors look like)
    return rng.standard_normal((1, P)) * rng.uniform(0, 2) + rng.standard_normal((n, P)) * (0.7 ** np.arange(n))[:, None]


def test_labels_identical_to_sklearn_on_600_random_grams():
    rng = np.random.default_rng(0)
    warnings.simplefilter("ignore")
    bad = []
    for trial in range(600):
        n = int(rng.integers(2, 33))
        k = int(rng.integers(2, min(4, n) + 1))
        T = _tasks(rng, trial % 3, n, k)
        F = clustering.embedding_from_gram(T @ T.T)
        ref = clustering.compute_kmeans_clustering_sklearn(F.copy(), k)
        got = clustering.compute_kmeans_clustering(F, k)
        if not np.array_equal(ref, got):
            bad.append((trial, n, k, ref.tolist(), got.tolist()))
    assert not bad, bad[:3]


def test_general_feature_matrices_and_other_seeds():
    rng = np.random.default_rng(5)
    from sklearn.cluster import KMeans
    for trial in range(60):
        n, d = int(rng.integers(3, 40)), int(rng.integers(1, 50))
        k = int(rng.integers(1, min(6, n) + 1))
        X = rng.standard_normal((n, d)).astype(np.float32) * rng.uniform(0.01, 100)
        seed = int(rng.integers(0, 2 ** 31 - 1))
        ref = KMeans(n_clusters=k, random_state=seed, n_init=10).fit_predict(X.copy())
        assert np.array_equal(ref, clustering.compute_kmeans_clustering(X, k, random_state=seed)), (trial, n, d, k)


def test_default_backend_reproduces_reference_partition_on_unclustered_tasks():
    """The case the exhaustive 'exact' backend got wrong (14/80): unclustered 8-task inputs of the
    bench's family.  Default backend vs the reference's full-feature k-means on the flattened task vectors."""
    rng = np.random.default_rng(11)
    names = [f"task{i:02d}" for i in range(8)]
    differ_exact = 0
    for trial in range(40):
        P = 600
        if trial % 2 == 0:      # iid
            X = rng.standard_normal((8, P)) * 0.01
        else:                   # decaying spectrum (synth "parity" family)
            A = np.linalg.qr(rng.standard_normal((8, 8)))[0]
            X = (A * (0.6 ** np.arange(8))) @ rng.standard_normal((8, P)) * 0.01 + 1e-4 * rng.standard_normal((8, P))
        tv = {n: {"w": torch.from_numpy(X[i].astype(np.float32))} for i, n in enumerate(names)}
        ref = R.cluster_tasks_full(tv, 2)
        X32 = X.astype(np.float32).astype(np.float64)
        G = X32 @ X32.T
        mine = clustering.cluster_from_gram(G, names, 2)
        assert all((ref[a] == ref[b]) == (mine[a] == mine[b]) for a in names for b in names), trial
        ex = clustering.cluster_from_gram(G, names, 2, backend="exact")
        differ_exact += not all((ref[a] == ref[b]) == (ex[a] == ex[b]) for a in names for b in names)
    print(f"'exact' backend differs from the reference partition in {differ_exact}/40 of these cases")


def test_invalid_k_raises_like_the_reference():
    with pytest.raises(ValueError):
        clustering.compute_kmeans_clustering(np.zeros((3, 3), np.float32), 4)
    with pytest.raises(ValueError):
        clustering.compute_kmeans_clustering(np.zeros((3, 3), np.float32), 0)


def test_duplicate_points_and_k_equals_n():
    from sklearn.cluster import KMeans
    warnings.simplefilter("ignore")
    X = np.asarray([[1, 0], [1, 0], [0, 1], [0, 1], [5, 5]], np.float32)
    for k in (2, 3, 5):
        ref = KMeans(n_clusters=k, random_state=42, n_init=10).fit_predict(X.copy())
        assert np.array_equal(ref, clustering.compute_kmeans_clustering(X, k)), k
