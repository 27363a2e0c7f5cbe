"""numpy restatement of the k-means / gap-statistic path.

TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference delegates the Lloyd loop and the pairwise distances to
scikit-learn (not vendored, version unpinned by the reference; the container
has scikit-learn 1.9.0 which is the de-facto pin).  The functions below restate
the published algorithm of that version, citing its files as
``sklearn/...:line``, and the reference's own call sites / definitions in
``p2_clustering_optK.py``.
"""
from __future__ import annotations

import numpy as np

__all__ = ["tolerance", "e_step", "lloyd_single", "kmeans_plusplus", "kmeans_fit",
           "kmeans_predict", "pairwise_euclidean", "inertia_v1", "inertia_v2",
           "elbow_distortion", "gap_statistic", "same_clustering"]


def tolerance(X, tol=1e-4):
    """sklearn/cluster/_kmeans.py:285-293: tol * mean(var(X, axis=0))."""
    return np.mean(np.var(X, axis=0)) * tol if tol != 0 else 0


def e_step(X, centers):
    """Labels by argmin_j (||c_j||^2 - 2 x.c_j), ||x||^2 omitted, strict '<' so the
    lowest index wins ties (sklearn/cluster/_k_means_lloyd.pyx:196-213)."""
    cn = np.sum(centers * centers, axis=1)
    d = cn[None, :] - 2.0 * (X @ centers.T)
    return np.argmin(d, axis=1).astype(np.int32)


def _m_step(X, labels, centers_old):
    """Sums, empty-cluster relocation, averaging, per-centre shift
    (_k_means_lloyd.pyx:118-165, _k_means_common.pyx:167-212,236-260)."""
    K, D = centers_old.shape
    sums = np.zeros((K, D), dtype=X.dtype)
    np.add.at(sums, labels, X)
    w = np.bincount(labels, minlength=K).astype(X.dtype)
    empty = np.where(w == 0)[0]
    if empty.size:
        dist = ((X - centers_old[labels]) ** 2).sum(axis=1)
        if np.max(dist) != 0:
            far = np.argpartition(dist, -empty.size)[:-empty.size - 1:-1]
            for new_id, idx in zip(empty, far):
                old_id = labels[idx]
                sums[old_id] -= X[idx]
                sums[new_id] = X[idx]
                w[new_id] = 1
                w[old_id] -= 1
    new = centers_old.copy()
    nz = w > 0
    new[nz] = sums[nz] / w[nz, None]
    shift = np.sqrt(((new - centers_old) ** 2).sum(axis=1))
    return new, shift


def _inertia(X, centers, labels):
    """_k_means_common.pyx:96-124: sum ||x - c_label||^2 computed directly."""
    return float(((X - centers[labels]) ** 2).sum(dtype=np.float64))


def lloyd_single(X, centers_init, max_iter=300, tol=0.0):
    """sklearn/cluster/_kmeans.py:630-757 (_kmeans_single_lloyd), dense path."""
    centers = np.array(centers_init, dtype=X.dtype, copy=True)
    labels_old = np.full(X.shape[0], -1, dtype=np.int32)
    strict = False
    n_iter = 0
    for i in range(max_iter):
        n_iter = i + 1
        labels = e_step(X, centers)
        centers, shift = _m_step(X, labels, centers)
        if np.array_equal(labels, labels_old):
            strict = True
            break
        if (shift ** 2).sum() <= tol:
            break
        labels_old = labels
    if not strict:
        labels = e_step(X, centers)
    return labels, _inertia(X, centers, labels), centers, n_iter


def kmeans_plusplus(X, n_clusters, rng):
    """sklearn/cluster/_kmeans.py:180-283 with n_local_trials = 2 + int(ln k).

    ``rng`` is a numpy RandomState (the reference leaves random_state=None so the
    global stream seeded at p2_clustering_optK.py:23 is consumed).
    """
    n = X.shape[0]
    trials = 2 + int(np.log(n_clusters))
    centers = np.empty((n_clusters, X.shape[1]), dtype=X.dtype)
    first = rng.choice(n, p=np.full(n, 1.0 / n))
    centers[0] = X[first]
    closest = ((X - centers[0]) ** 2).sum(axis=1)
    pot = closest.sum()
    for c in range(1, n_clusters):
        rv = rng.uniform(size=trials) * pot
        cand = np.searchsorted(np.cumsum(closest), rv)
        np.clip(cand, None, n - 1, out=cand)
        dc = ((X[cand][:, None, :] - X[None, :, :]) ** 2).sum(axis=2)
        np.minimum(closest, dc, out=dc)
        pots = dc.sum(axis=1)
        best = int(np.argmin(pots))
        pot, closest = pots[best], dc[best]
        centers[c] = X[cand[best]]
    return centers


def same_clustering(a, b, K):
    """_k_means_common.pyx:_is_same_clustering: equal up to a label permutation."""
    mapping = np.full(K, -1, dtype=np.int64)
    for x, y in zip(a, b):
        if mapping[x] == -1:
            mapping[x] = y
        elif mapping[x] != y:
            return False
    return True


def kmeans_fit(X, n_clusters, init="k-means++", n_init=1, max_iter=300, tol=1e-4, rng=None):
    """KMeans.fit, sklearn/cluster/_kmeans.py:1463-1560: mean-centre, best of n_init by
    inertia, centres shifted back.  ``init`` is 'k-means++' or a (K, D) array.

    Returns dict(labels, centers, inertia, n_iter); ``labels`` come from the centred
    run exactly as ``fit_predict`` returns them.
    """
    X = np.ascontiguousarray(X)
    mean = X.mean(axis=0)
    Xc = X - mean
    tol_eff = tolerance(Xc, tol)
    best = None
    for _ in range(n_init):
        if isinstance(init, str):
            c0 = kmeans_plusplus(Xc, n_clusters, rng)
        else:
            c0 = np.asarray(init, dtype=X.dtype) - mean
        labels, inertia, centers, n_iter = lloyd_single(Xc, c0, max_iter, tol_eff)
        if best is None or (inertia < best["inertia"]
                            and not same_clustering(labels, best["labels"], n_clusters)):
            best = dict(labels=labels, centers=centers, inertia=inertia, n_iter=n_iter)
    best["centers"] = best["centers"] + mean
    return best


def kmeans_predict(X, centers):
    """KMeans.predict: E-step on the un-centred data with the shifted-back centres."""
    return e_step(np.ascontiguousarray(X), np.asarray(centers, dtype=X.dtype))


def pairwise_euclidean(X, chunk=2048):
    """sklearn.metrics.pairwise_distances(X) (euclidean), sklearn/metrics/pairwise.py:
    376-427 and 567-640: float32 input is up-cast to float64 per chunk for
    ||x||^2 + ||y||^2 - 2xy, stored back as float32, clamped at 0, the diagonal zeroed,
    then sqrt.  float64 input is computed directly in float64.
    """
    n = X.shape[0]
    out = np.empty((n, n), dtype=X.dtype)
    X64 = X.astype(np.float64)
    nn = (X64 * X64).sum(axis=1)
    for i0 in range(0, n, chunk):
        d = -2.0 * (X64[i0:i0 + chunk] @ X64.T)
        d += nn[i0:i0 + chunk, None]
        d += nn[None, :]
        out[i0:i0 + chunk] = d.astype(X.dtype)
    np.maximum(out, 0, out=out)
    np.fill_diagonal(out, 0)
    return np.sqrt(out, out=out)


def inertia_v1(a, X):
    """KM.compute_inertia_v1, p2_clustering_optK.py:334-342: mean over clusters of the
    mean of the full n_c x n_c Euclidean distance matrix (zero diagonal included)."""
    return np.mean([np.mean(pairwise_euclidean(X[a == c])) for c in np.unique(a)])


def inertia_v2(a, X):
    """KM.computer_intertia_v2 (sic), p2_clustering_optK.py:344-351:
    sum_c (sum of the full distance matrix) / (2 n_c)."""
    wk = 0
    for c in np.unique(a):
        wk = wk + np.sum(pairwise_euclidean(X[a == c])) / (2 * (a == c).sum())
    return wk


def elbow_distortion(X, centers):
    """p2_clustering_optK.py:258-265: sum_i min_j ||x_i - c_j|| / N (cdist, float64)."""
    X = np.asarray(X, dtype=np.float64)
    C = np.asarray(centers, dtype=np.float64)
    d = np.sqrt(((X[:, None, :] - C[None, :, :]) ** 2).sum(axis=2))
    return float(d.min(axis=1).sum() / X.shape[0])


def gap_statistic(fit_predict, data, k_max=5, n_references=5, version=1, draw=None):
    """KM.compute_gap_internal_metric, p2_clustering_optK.py:353-410, without the
    internal-metric columns.

    ``fit_predict(k, X) -> labels`` stands in for the sklearn-like object whose
    ``n_clusters`` the reference sets at :367; ``draw(shape)`` stands in for
    ``np.random.random_sample`` (:370).  Returns dict of per-k arrays
    (k, gap, ref, act, ref_s).
    """
    if draw is None:
        draw = np.random.random_sample
    if data.ndim == 1:
        data = data.reshape(-1, 1)
    inertia = inertia_v1 if version == 1 else inertia_v2
    rng_ = data.max() - data.min()                                  # :360
    rows = []
    for k in range(2, k_max + 1):
        local = []
        for _ in range(n_references):
            reference = draw(data.shape) * rng_ + data.min()        # :370 (float64)
            local.append(inertia(fit_predict(k, reference), reference))
        ref = np.mean(np.log(local))                                # :374
        ref_s = np.sqrt(1 + 1 / n_references) * np.std(np.log(local))   # :375
        act = np.log(inertia(fit_predict(k, data), data))           # :377-379
        rows.append((k, ref - act, ref, act, ref_s))
    cols = list(zip(*rows))
    return {n: np.array(c) for n, c in zip(("k", "gap", "ref", "act", "ref_s"), cols)}


# ---- cluster-validity metrics of the gap loop (internal_eval.py) -------------------------------------------------
def dunn_index(x, labels):
    """DunnIndex.__call__, internal_eval.py:83-109: nearest-point distance of every pair of clusters (:37-53; the
    K x K matrix starts at +inf off the diagonal), zero entries dropped through ``nonzero()`` (:106), over the largest
    "farthest" diameter (:77-81).  Vectorised restatement of the reference's pure-Python double loops; distances are
    sklearn's euclidean_distances (float32 input evaluated through the float64 Gram form, :103)."""
    x = np.asarray(x)
    labels = np.asarray(labels)
    d = pairwise_euclidean(x)
    ks = np.unique(labels)
    ic = []
    diam = 0.0
    for i, a in enumerate(ks):
        ma = labels == a
        if ma.sum() > 1:
            diam = max(diam, float(d[np.ix_(ma, ma)].max()))
        for b in ks[i + 1:]:
            ic.append(float(d[np.ix_(ma, labels == b)].min()))
    ic = np.array([v for v in ic if v != 0.0])
    if ic.size == 0:
        raise ValueError("min() iterable argument is empty")
    return float(ic.min() / diam)


def calinski_harabasz(x, labels):
    """CHIndex, internal_eval.py:125-135 -> sklearn.metrics.calinski_harabasz_score (_unsupervised.py): float64."""
    x = np.asarray(x, np.float64)
    labels = np.asarray(labels)
    ks = np.unique(labels)
    mean = x.mean(0)
    extra = intra = 0.0
    for k in ks:
        xk = x[labels == k]
        ck = xk.mean(0)
        extra += len(xk) * ((ck - mean) ** 2).sum()
        intra += ((xk - ck) ** 2).sum()
    n, K = x.shape[0], len(ks)
    return 1.0 if intra == 0.0 else float(extra * (n - K) / (intra * (K - 1.0)))


def davies_bouldin(x, labels):
    """DBIndex, internal_eval.py:138-147 -> sklearn.metrics.davies_bouldin_score."""
    x = np.asarray(x, np.float64)
    labels = np.asarray(labels)
    ks = np.unique(labels)
    cen = np.stack([x[labels == k].mean(0) for k in ks])
    s = np.array([np.sqrt(((x[labels == k] - cen[i]) ** 2).sum(1)).mean() for i, k in enumerate(ks)])
    cd = np.sqrt(((cen[:, None] - cen[None]) ** 2).sum(2))
    if np.allclose(s, 0) or np.allclose(cd, 0):
        return 0.0
    cd[cd == 0] = np.inf
    return float(np.max((s[:, None] + s[None, :]) / cd, axis=1).mean())
