"""GPU parity of the k-means / gap-statistic path against scikit-learn 1.9.0 results
(tests/golden/kmeans.npz, gap.npz - produced by oracle/gen_golden.py through sklearn and the
reference's own KM class).

Labels are compared bit-exactly; a mismatch is tolerated only where the float64 margin between
the two nearest centres is below 1e-5 of the distance (a documented near-tie, BASELINE.json).
"""
import numpy as np
import pytest
import torch

from gpu_util import record

pytestmark = pytest.mark.gpu


def _labels_equal_mod_ties(name, got, want, X, centers, tie=1e-5, max_frac=1.0):
    got, want = np.asarray(got), np.asarray(want)
    bad = np.nonzero(got != want)[0]
    if bad.size == 0:
        return
    assert bad.size <= max_frac * got.size, f"{name}: {bad.size} label mismatches of {got.size}"
    X64, C64 = np.asarray(X, np.float64), np.asarray(centers, np.float64)
    d = ((X64[bad, None, :] - C64[None]) ** 2).sum(2)
    d.sort(axis=1)
    margin = (d[:, 1] - d[:, 0]) / np.maximum(d[:, 0], 1e-30)
    assert np.all(margin < tie), f"{name}: {bad.size} label mismatches, worst margin {margin.max():.3e}"


def _centres_close_mod_flips(name, km, ref, X, rtol=1e-5, init=None, tie=5e-5):
    """Centres after several Lloyd iterations, with the near-tie rows accounted for: a row that sklearn's float32 GEMM
    and the kernel assign to different clusters (a documented near-tie, checked by _labels_equal_mod_ties) moves the two
    centres involved by (x - c) / n_k.  The per-cluster bound is therefore
        1e-5 (1 + |c|)  +  2 * flips_k * max_i |x_i - c_k|_inf / n_k
    with flips_k counted on the final labels (the E-steps inside the loop are not observable from outside; the factor 2
    covers them).  Without flips this is the plain 1e-5 check."""
    X64 = np.asarray(X, np.float64)
    got, want = np.asarray(km.cluster_centers_, np.float64), np.asarray(ref.cluster_centers_, np.float64)
    bad = km.labels_ != ref.labels_
    if init is not None:      # one-iteration fits: the rows the E-step on `init` could have decided either way
        d = ((X64[:, None, :] - np.asarray(init, np.float64)[None]) ** 2).sum(2)
        d.sort(axis=1)
        bad = bad | ((d[:, 1] - d[:, 0]) < tie * np.maximum(d[:, 0], 1e-30))
    worst = 0.0
    for k in range(want.shape[0]):
        members = ref.labels_ == k
        n_k = max(int(members.sum()), 1)
        flips = int((bad & (members | (km.labels_ == k))).sum())
        spread = float(np.abs(X64[members] - want[k]).max()) if members.any() else 0.0
        lim = rtol * (1.0 + np.abs(want[k])) + 2.0 * flips * spread / n_k
        worst = max(worst, float((np.abs(got[k] - want[k]) / lim).max()))
    record(name, worst, 0.0, 0.0, 1.0)        # logged as a ratio to the bound (<= 1 passes)


@pytest.mark.parametrize("dtag", ["f32", "f64"])
@pytest.mark.parametrize("k", [2, 4, 7])
def test_lloyd_matches_sklearn(golden, dtag, k):
    from deep_interpolation_clustering_b200.kmeans import KMeansB200
    g = golden("kmeans")
    X = g["X"].astype(np.float32 if dtag == "f32" else np.float64)
    pre = f"{dtag}_k{k}_"
    km = KMeansB200(n_clusters=k, init=X[:k].copy(), n_init=1).fit(X)
    assert km.labels_.dtype == np.int32 and km.cluster_centers_.dtype == X.dtype
    _labels_equal_mod_ties(pre + "labels", km.labels_, g[pre + "labels"], X, g[pre + "centers"])
    assert km.n_iter_ == int(g[pre + "n_iter"])
    record(pre + "centers", km.cluster_centers_, g[pre + "centers"], 1e-5, 1e-5)
    record(pre + "inertia", km.inertia_, g[pre + "inertia"], 1e-5, 0)
    Xv = g["Xv"].astype(X.dtype)
    _labels_equal_mod_ties(pre + "predict", km.predict(Xv), g[pre + "predict"], Xv, g[pre + "centers"])
    record(pre + "elbow", km.score_distortion(X), g[pre + "elbow_train"], 1e-5, 0)
    # fit_predict is fit().labels_, and cluster_centers_ is a writable ndarray (p4 mutates it)
    km.cluster_centers_[0] += 0.0


def test_empty_cluster_relocation(golden):
    from deep_interpolation_clustering_b200.kmeans import KMeansB200
    g = golden("kmeans")
    km = KMeansB200(n_clusters=3, init=g["reloc_init"], n_init=1).fit(g["X"])
    _labels_equal_mod_ties("reloc", km.labels_, g["reloc_labels"], g["X"], g["reloc_centers"])
    assert km.n_iter_ == int(g["reloc_n_iter"])
    record("reloc_centers", km.cluster_centers_, g["reloc_centers"], 1e-5, 1e-5)
    record("reloc_inertia", km.inertia_, g["reloc_inertia"], 1e-5, 0)


def test_inertia_definitions(golden):
    from deep_interpolation_clustering_b200.gap import KM
    g = golden("kmeans")
    a, X = g["f32_k4_labels"], g["X"]
    km = KM(5)
    record("inertia_v1_f32", km.compute_inertia_v1(a, X), g["inertia_v1_f32"], 1e-5, 0)
    record("inertia_v2_f32", km.computer_intertia_v2(a, X), g["inertia_v2_f32"], 1e-5, 0)
    kx = KM(5, exact_pairwise=True)             # direct float64 kernel (the default goes to the tensor cores)
    record("inertia_v1_f64", kx.compute_inertia_v1(a, X.astype(np.float64)), g["inertia_v1_f64"], 1e-9, 0)
    record("inertia_v2_f64", kx.computer_intertia_v2(a, X.astype(np.float64)), g["inertia_v2_f64"], 1e-9, 0)
    record("inertia_v1_f64_tc", km.compute_inertia_v1(a, X.astype(np.float64)), g["inertia_v1_f64"], 1e-6, 0)


class _FixedInit:
    """Same duck type gen_golden.py handed to the reference: init = first k rows."""
    accepts_device_tensors = False

    def __init__(self):
        self.n_clusters = 2

    def fit_predict(self, X):
        from deep_interpolation_clustering_b200.kmeans import KMeansB200
        X = np.ascontiguousarray(X)
        return KMeansB200(n_clusters=self.n_clusters, init=X[:self.n_clusters].copy(), n_init=1).fit_predict(X)


@pytest.mark.parametrize("version", [1, 2])
def test_gap_statistic_dataframe(golden, version):
    from deep_interpolation_clustering_b200.gap import KM
    g = golden("gap")
    names = ["Sihouette", "Davies-Bouldin_Index", "Calinski-Harabasz"]
    km = KM(5, None, names, 1, 3)
    np.random.seed(7)
    df = km.compute_gap_internal_metric(_FixedInit(), g["X"], k_max=5, n_references=3, version=version)
    assert list(df.columns) == ["k", "gap", "ref", "act", "ref_s"] + names
    assert list(df.index) == [2, 3, 4, 5]
    for col in ("k", "gap", "ref", "act"):
        record(f"gap_v{version}/{col}", df[col].to_numpy(np.float64), g[f"v{version}_{col}"], 1e-5, 1e-6)
    # ref_s = std over the reference sets of log(inertia) * sqrt(1 + 1/B) (p2_clustering_optK.py:391-393): a difference of
    # the nearly equal log-inertias that are themselves held to 1e-5 above.  Their spread here is ~1e-2 of their size, so
    # a 1e-5 perturbation of each is 1e-3 of the standard deviation: the tolerance is the conditioning of the statistic,
    # not a looser kernel (the float32 pairwise sums on the fixture agree to 2e-7).
    record(f"gap_v{version}/ref_s", df["ref_s"].to_numpy(np.float64), g[f"v{version}_ref_s"], 1e-3, 1e-6)
    for col in names:
        record(f"gap_v{version}/{col}", df[col].to_numpy(np.float64), g[f"v{version}_{col}"], 1e-5, 1e-6)


def test_kmeans_plusplus_quality_and_stream(golden):
    """k-means++ seeding on device: reaches the blob optimum, and consumes the numpy stream in
    the same quantity as sklearn (1 + (K-1) * (2 + int(ln K)) doubles per init)."""
    from sklearn.cluster import KMeans
    from deep_interpolation_clustering_b200.kmeans import KMeansB200
    g = golden("kmeans")
    X = g["X"]
    ref = KMeans(n_clusters=5, n_init=3, random_state=0).fit(X)
    km = KMeansB200(n_clusters=5, n_init=3, random_state=0).fit(X)
    assert km.inertia_ <= ref.inertia_ * 1.02
    r1, r2 = np.random.RandomState(5), np.random.RandomState(5)
    KMeans(n_clusters=5, n_init=2, random_state=r1).fit(X)
    KMeansB200(n_clusters=5, n_init=2, random_state=r2).fit(X)
    assert r1.uniform() == r2.uniform()


def test_config4_size_properties():
    """1M x 64 (BASELINE config 4 shape) run to strict convergence (tol = 0): fixed point
    (predict == labels_ away from ties), centres are the means of their members, inertia equals
    the direct sum."""
    from deep_interpolation_clustering_b200.kmeans import KMeansB200
    from deep_interpolation_clustering_b200 import synth
    dev = torch.device("cuda:0")
    X = torch.from_numpy(synth.make_blobs(1_000_000, 64, 5, seed=4)).to(dev)
    km = KMeansB200(n_clusters=5, n_init=1, random_state=3, tol=0.0).fit(X)
    lab, cen = km.labels_.long(), km.cluster_centers_
    means = torch.zeros_like(cen, dtype=torch.float64).index_add_(0, lab, X.double())
    means /= torch.bincount(lab, minlength=5).double()[:, None]
    assert torch.allclose(means.float(), cen, rtol=1e-4, atol=1e-4)
    direct = float(((X.double() - cen.double()[lab]) ** 2).sum())
    assert abs(direct - km.inertia_) <= 1e-5 * direct
    again = km.predict(X).long()
    frac = float((again != lab).double().mean())
    assert frac < 1e-3, frac


@pytest.mark.parametrize("draw", ["device", "device32"])
def test_gap_sweep_device_draws_and_task_mode(draw):
    """Reference sets drawn on the GPU (float64 like the reference's np.random draws, or float32 opt-in): the gap
    statistic flattens at the number of blobs, the per-task seeded sweep is reproducible, and float32 draws move the
    reference term by no more than its own spread."""
    from deep_interpolation_clustering_b200 import synth
    from deep_interpolation_clustering_b200.gap import KM
    from deep_interpolation_clustering_b200.kmeans import KMeansB200
    X = synth.make_blobs(6000, 64, 4, seed=12).astype(np.float32)
    def sweep(d):
        return KM(6).compute_gap_internal_metric(KMeansB200(n_init=2, random_state=2), X, k_max=6, n_references=3,
                                                 version=1, draw=d, task_parallel=True, seed=1).astype(float)
    a, b = sweep(draw), sweep(draw)
    assert np.array_equal(a.to_numpy(), b.to_numpy())
    assert a["gap"][4] - a["gap"][3] > 10 * abs(a["gap"][5] - a["gap"][4])      # the gap flattens at the 4 blobs
    ref64 = sweep("device")
    assert np.allclose(a["ref"], ref64["ref"], atol=5e-3)
    assert np.allclose(a["act"], ref64["act"], rtol=1e-6)


def test_errors():
    from deep_interpolation_clustering_b200.kmeans import KMeansB200
    with pytest.raises(ValueError):
        KMeansB200(n_clusters=5).fit(np.zeros((3, 4), np.float32))
    with pytest.raises(ValueError):
        KMeansB200(n_clusters=2, init=np.zeros((3, 4))).fit(np.zeros((10, 4), np.float32))


@pytest.mark.parametrize("n,D", [(3000, 64), (1000, 40), (700, 256), (4097, 8), (16385, 64), (40000, 48), (513, 4),
                                 (16385, 256), (9000, 128), (5001, 100), (2000, 192), (3000, 300)])
def test_tensor_core_pairwise_matches_exact(n, D):
    """tcgen05 (3xTF32) pairwise-distance sum vs the direct float64 kernel and numpy."""
    from deep_interpolation_clustering_b200 import synth
    from deep_interpolation_clustering_b200.gap import pairwise_dist_sum
    X = synth.make_blobs(n, D, 3, seed=n)
    Xd = torch.from_numpy(X).cuda()
    tc = float(pairwise_dist_sum(Xd))
    exact = float(pairwise_dist_sum(Xd.double(), exact=True))
    assert np.isfinite(tc), "tensor-core kernel reported a timeout (NaN)"
    record(f"pairwise_tc_n{n}_D{D}", tc, exact, 2e-6, 0)
    if n <= 3000:
        X64 = X.astype(np.float64)
        ref = np.sqrt(np.maximum(((X64[:, None, :] - X64[None]) ** 2).sum(2), 0)).sum()
        record(f"pairwise_exact_n{n}_D{D}", exact, ref, 1e-10, 0)


@pytest.mark.parametrize("n,D,dtype", [(600, 64, "f32"), (16385, 64, "f32"), (5000, 32, "f32"), (3000, 256, "f32"),
                                       (700, 100, "f32"), (900, 64, "f64"), (300, 16, "f32")])
@pytest.mark.parametrize("n_parts", [2, 3, 8])
def test_pairwise_stripes_add_up(n, D, dtype, n_parts):
    """dic_pairwise_dist_sum_part: the stripes of the tile list (one per rank of a row-sharded sweep) tile it exactly
    once on every kernel - TMA-fed tcgen05 (D <= 64), register-staged tcgen05 (D > 64), CUDA cores (float64, odd D,
    small n) - including stripes that own no tile at all (n = 300 / 600 with 8 parts)."""
    from deep_interpolation_clustering_b200 import synth
    from deep_interpolation_clustering_b200.gap import pairwise_dist_sum, pairwise_dist_sum_part
    X = synth.make_blobs(n, D, 3, seed=n + D)
    Xd = torch.from_numpy(X).cuda().to(torch.float32 if dtype == "f32" else torch.float64)
    full = float(pairwise_dist_sum(Xd))
    parts = [float(pairwise_dist_sum_part(Xd, r, n_parts)) for r in range(n_parts)]
    assert all(np.isfinite(v) and v >= 0 for v in parts), parts
    if n >= 3000:
        assert min(parts) > 0.5 * full / n_parts, "unbalanced stripes"
    record(f"pairwise_parts_n{n}_D{D}_{dtype}_p{n_parts}", sum(parts), full, 1e-9, 0)


@pytest.mark.parametrize("n,D,K", [(5000, 64, 5), (3001, 20, 7), (20000, 32, 3), (4000, 256, 4), (2500, 128, 3),
                                   (3000, 100, 5)])
def test_silhouette_tensor_core_matches_float64(n, D, K):
    """dic_cluster_rowsums (tcgen05 row sums by cluster) vs the chunked float64 distance path and sklearn:
    ragged cluster sizes, a singleton cluster, D < 64."""
    from deep_interpolation_clustering_b200 import synth
    from deep_interpolation_clustering_b200.internal_eval import Sihouette
    X = synth.make_blobs(n, D, K, seed=n + D)
    rng = np.random.RandomState(n)
    labels = rng.randint(0, K, size=n)
    labels[:n // 2] = np.argmin(((X[:n // 2, None, :] - X[None, :K, :]) ** 2).sum(2), axis=1)   # some structure
    labels[labels == K - 1] = 0
    labels[7] = K - 1                                                                          # singleton cluster
    Xd = torch.from_numpy(X).cuda()
    native = Sihouette(native=True)(Xd, labels)
    chunked = Sihouette(native=False)(Xd, labels)
    record(f"silhouette_tc_n{n}_D{D}_K{K}", native, chunked, 1e-5, 1e-7)
    # the (N, K) sums themselves
    inv = torch.from_numpy(np.unique(labels, return_inverse=True)[1]).cuda()
    Kp = int(inv.max()) + 1
    sums = Sihouette.rowsums_native(Xd, inv, Kp).cpu().numpy()
    rows = rng.choice(n, size=64, replace=False)
    X64 = X.astype(np.float64)
    inv_h = inv.cpu().numpy()
    ref = np.zeros((64, Kp))
    for a, i in enumerate(rows):
        d = np.sqrt(((X64[i] - X64) ** 2).sum(1))
        ref[a] = np.bincount(inv_h, weights=d, minlength=Kp)
    # float32-grade sums (split operands, float32 accumulation in the tensor core, sqrt.approx).  Intra-blob distances
    # are differences of large norms (|x|^2 ~ 17 D against d^2 ~ 2 D here) and the accumulator's error grows with the
    # number of products, so the raw sums get 3e-5 above D = 64; the silhouette itself stays within 1e-5 (above).
    record(f"cluster_rowsums_n{n}_D{D}_K{K}", sums[rows], ref, 1e-5 if D <= 64 else 3e-5, 1e-6)
    if n <= 5000:
        from sklearn.metrics import silhouette_score
        record(f"silhouette_vs_sklearn_n{n}", native, float(silhouette_score(X, labels)), 1e-5, 1e-6)


@pytest.mark.parametrize("dtag,D,K", [("f32", 64, 3), ("f32", 64, 10), ("f32", 64, 16), ("f64", 64, 4), ("f64", 64, 10),
                                      ("f32", 128, 5), ("f64", 32, 6), ("f32", 100, 7), ("f64", 100, 3),
                                      ("f32", 200, 4), ("f32", 300, 5), ("f32", 64, 20), ("f32", 256, 4), ("f32", 256, 10),
                                      ("f32", 256, 16), ("f64", 128, 6), ("f64", 128, 12)])
def test_lloyd_kernel_dispatch_matches_sklearn(dtag, D, K):
    """Every Lloyd kernel (specialised tile kernel: rows of 16 / 32 vectors; streaming row kernel; generic tile
    kernels) against scikit-learn run here on the same data and the same initial centres.  The comparison is made
    after ONE (and, for few clusters, THREE) iterations: with more clusters than blobs the converged solution is
    chaotic in the last bits of early near-ties (and sklearn's own sums depend on its thread count), which says
    nothing about a kernel."""
    import warnings
    from sklearn.cluster import KMeans
    from deep_interpolation_clustering_b200 import synth
    from deep_interpolation_clustering_b200.kmeans import KMeansB200
    dt = np.float32 if dtag == "f32" else np.float64
    X = synth.make_blobs(3001, D, 6, seed=D + K).astype(dt)
    init = X[:K].copy()
    Xv = synth.make_blobs(777, D, 6, seed=D + K + 1).astype(dt)
    # three iterations only where clusters <= blobs: beyond that ONE flipped near-tie moves a small cluster's centre
    # by ~1/n_c and legitimately re-labels borderline rows two iterations later (seen at K = 16: 4 of 3001)
    for iters in ((1, 3) if K <= 6 else (1,)):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")                      # ConvergenceWarning: max_iter is the point here
            ref = KMeans(n_clusters=K, init=init, n_init=1, max_iter=iters, tol=0.0).fit(X)
        km = KMeansB200(n_clusters=K, init=init, n_init=1, max_iter=iters, tol=0.0).fit(X)
        tag = f"dispatch_{dtag}_D{D}_K{K}_it{iters}"
        _labels_equal_mod_ties(tag + "_labels", km.labels_, ref.labels_, X, ref.cluster_centers_)
        assert km.n_iter_ == ref.n_iter_
        record(tag + "_centers", km.cluster_centers_, ref.cluster_centers_, 1e-5, 1e-5)
        record(tag + "_inertia", km.inertia_, ref.inertia_, 1e-5, 0)
        _labels_equal_mod_ties(tag + "_predict", km.predict(Xv), ref.predict(Xv), Xv, ref.cluster_centers_)


def test_exact_pairwise_is_honoured_for_float32_clusters():
    """KM(exact_pairwise=True) / pairwise_dist_sum(exact=True): float32 clusters of >= 512 rows must take the direct
    (x_i - x_j)^2 kernel (dtype | DIC_PAIRWISE_EXACT), not the Gram-form tensor-core kernel on UNCENTRED rows.  With a
    mean offset of 1e3 the Gram form in float32 loses the distances entirely; the direct form does not care."""
    from deep_interpolation_clustering_b200 import synth
    from deep_interpolation_clustering_b200.gap import KM, pairwise_dist_sum
    X = (synth.make_blobs(2000, 64, 3, seed=11) + 1000.0).astype(np.float32)
    Xd = torch.from_numpy(X).cuda()
    truth = float(pairwise_dist_sum(Xd.double(), exact=True))
    X64 = X.astype(np.float64)
    ref = np.sqrt(np.maximum(((X64[:, None, :] - X64[None]) ** 2).sum(2), 0)).sum()
    record("pairwise_exact_f64_offset", truth, ref, 1e-10, 0)
    record("pairwise_exact_f32_offset", float(pairwise_dist_sum(Xd, exact=True)), truth, 2e-6, 0)
    record("pairwise_default_f32_offset", float(pairwise_dist_sum(Xd)), truth, 2e-6, 0)     # centred by the wrapper
    a = np.arange(2000) % 2
    v_exact = KM(5, exact_pairwise=True).compute_inertia_v1(a, X)
    v_ref = np.mean([np.sqrt(np.maximum(((X64[a == c][:, None] - X64[a == c][None]) ** 2).sum(2), 0)).mean()
                     for c in (0, 1)])
    record("inertia_v1_exact_f32_offset", v_exact, v_ref, 2e-6, 0)


def test_stalled_pipeline_sentinel_raises():
    from deep_interpolation_clustering_b200 import _lib
    from deep_interpolation_clustering_b200.gap import check_pairwise_sums
    assert check_pairwise_sums([1.0, 0.0]) == [1.0, 0.0]
    with pytest.raises(_lib.DicError, match="stalled"):
        check_pairwise_sums([1.0, float("nan")])


@pytest.mark.parametrize("dtag,D,K", [("f32", 64, 4), ("f32", 64, 10), ("f64", 64, 4), ("f64", 64, 10), ("f64", 64, 16),
                                      ("f64", 64, 1), ("f32", 256, 10), ("f32", 128, 16),
                                      ("f32", 100, 7), ("f64", 32, 12), ("f32", 64, 16), ("f32", 256, 4), ("f32", 128, 2),
                                      ("f32", 256, 16), ("f32", 64, 1)])
@pytest.mark.parametrize("n", [20011, 128, 77])
def test_every_lloyd_kernel_through_the_abi_selector(dtag, D, K, n):
    """DIC_KM_KERNEL(k) in the flags addresses each Lloyd kernel explicitly (no environment switches): all of them -
    CUDA-core kernels 1-4, the tcgen05 E-step kernel 5 (float32) and its float64 form 6 - must produce the labels, sums, counts and changed-label
    count of the general kernel in the Lloyd-loop form of the pass; a kernel that does not cover the shape says so."""
    from deep_interpolation_clustering_b200 import synth
    from deep_interpolation_clustering_b200.kmeans import _Device
    dt = torch.float32 if dtag == "f32" else torch.float64
    X = torch.from_numpy(synth.make_blobs(n, D, 6, seed=D + K)).cuda().to(dt)
    cen = X[:K].clone().contiguous()
    prev = torch.from_numpy(np.random.RandomState(n).randint(0, K, size=n).astype(np.int32)).cuda()
    out = {}
    base = 1 | 4                                     # DIC_KM_COUNT_CHANGES | DIC_KM_NO_INERTIA: what a Lloyd iteration passes
    for sel in (4, 0, 1, 2, 3, 5, 6):
        st = _Device(X, K)
        st.labels.copy_(prev)
        try:
            st.assign(cen, base | sel << 8)
        except ValueError as e:
            assert "does not cover" in str(e) or "covers" in str(e) or "float32 only" in str(e) or \
                "float64 only" in str(e), str(e)
            continue
        out[sel] = (st.labels.cpu().numpy().copy(), st.sums.cpu().numpy().copy(), st.counts.cpu().numpy().copy(),
                    st.stats.cpu().numpy().copy())
    assert 4 in out and 0 in out and len(out) >= 3
    if dtag == "f32" and D in (64, 128, 256):
        assert 5 in out, "the tensor-core pass must cover this shape"
    if dtag == "f64" and D == 64:
        assert 6 in out, "the float64 tensor-core pass must cover this shape"
    lab, sums, counts, stats = out[4]
    for sel, (l2, s2, c2, t2) in out.items():
        _labels_equal_mod_ties(f"selector{sel}", l2, lab, X.cpu().numpy(), cen.cpu().numpy())
        assert np.isfinite(s2).all() and np.isfinite(t2).all(), f"kernel {sel}: stalled pipeline sentinel"
        assert c2.sum() == n
        if np.array_equal(l2, lab):
            # float32 kernels add a tile's rows in float32 before the float64 partials: 1e-6-grade agreement
            np.testing.assert_allclose(s2, sums, rtol=2e-6, atol=2e-6 * float(np.abs(sums).max()))
            np.testing.assert_array_equal(c2, counts)
            assert t2[1] == stats[1], f"kernel {sel}: changed-label count {t2[1]} vs {stats[1]}"
    with pytest.raises(ValueError):
        _Device(X, K).assign(cen, 9 << 8)


@pytest.mark.parametrize("dtag,D,K,sel", [("f32", 64, 10, 5), ("f32", 128, 7, 5), ("f32", 256, 4, 5), ("f64", 64, 10, 6),
                                          ("f64", 64, 16, 6)])
@pytest.mark.parametrize("want_sums", [True, False])
def test_tensor_core_lloyd_pass_many_tiles_per_cta(dtag, D, K, sel, want_sums):
    """The tcgen05 passes at a size where every CTA walks its unit ring around several times (300k rows = 16 tiles per
    CTA; the selector test above has at most two), in the Lloyd form (sums) and the predict form (labels only: the raw
    units are recycled as soon as the operand halves are taken), against the general CUDA-core kernel."""
    from deep_interpolation_clustering_b200 import synth
    from deep_interpolation_clustering_b200.kmeans import _Device
    n = 300_017
    dt = torch.float32 if dtag == "f32" else torch.float64
    X = torch.from_numpy(synth.make_blobs(n, D, 6, seed=D + K)).cuda().to(dt)
    cen = X[:K].clone().contiguous()
    prev = torch.from_numpy(np.random.RandomState(n).randint(0, K, size=n).astype(np.int32)).cuda()
    out = {}
    for which in (4, sel):
        st = _Device(X, K)
        st.labels.copy_(prev)
        for _ in range(2):                           # twice: the second launch must not depend on leftovers of the first
            st.labels.copy_(prev)
            st.assign(cen, 1 | 4 | which << 8, want_sums=want_sums)
        out[which] = (st.labels.cpu().numpy().copy(), st.sums.cpu().numpy().copy(), st.counts.cpu().numpy().copy(),
                      st.stats.cpu().numpy().copy())
    lab, sums, counts, stats = out[4]
    l2, s2, c2, t2 = out[sel]
    assert np.isfinite(t2).all(), "stalled pipeline sentinel"
    _labels_equal_mod_ties(f"many_tiles_{dtag}_D{D}_K{K}", l2, lab, X.cpu().numpy(), cen.cpu().numpy())
    if want_sums and np.array_equal(l2, lab):
        np.testing.assert_allclose(s2, sums, rtol=2e-6, atol=2e-6 * float(np.abs(sums).max()))
        np.testing.assert_array_equal(c2, counts)
        assert t2[1] == stats[1]


def test_tensor_core_lloyd_pass_inside_a_fit_matches_sklearn(monkeypatch):
    """A whole fit with every Lloyd iteration forced onto the tcgen05 pass (DIC_KM_KERNEL(5)) against scikit-learn with
    the same initial centres: labels mod near-ties, centres, inertia, iteration count (D = 64 and the real latent 256)."""
    import warnings
    from sklearn.cluster import KMeans
    from deep_interpolation_clustering_b200 import synth
    from deep_interpolation_clustering_b200 import kmeans as km_mod
    orig = km_mod._Device.lloyd_run

    def forced(self, centers, flags, n_steps, tol):
        return orig(self, centers, flags | 5 << 8, n_steps, tol)
    monkeypatch.setattr(km_mod._Device, "lloyd_run", forced)
    for D, K in ((64, 10), (256, 4), (128, 7)):
        X = synth.make_blobs(30011, D, 5, seed=D)
        cur = X[:K].copy()
        # Three Lloyd iterations, each compared on its own: iteration i starts from scikit-learn's centres after i - 1
        # iterations on BOTH sides.  (Comparing the end of a 3-iteration run instead lets one near-tie row of the first
        # E-step move two centres by |x - c| / n_k ~ 3e-3, after which dozens of rows sit inside the rounding band of
        # the later E-steps: a property of Lloyd's iteration, not of the kernel.)
        for it in (1, 2, 3):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                ref = KMeans(n_clusters=K, init=cur, n_init=1, max_iter=1, tol=0.0).fit(X)
            km = km_mod.KMeansB200(n_clusters=K, init=cur, n_init=1, max_iter=1, tol=0.0).fit(X)
            tag = f"tc_fit_D{D}_K{K}_it{it}"
            # float32-grade dots on both sides (sklearn: a float32 GEMM): rows whose float64 margin is below a few 1e-5
            # are decided by rounding.  labels_ of a one-iteration fit are the E-step on the NEW centres, and those
            # already carry the near-tie rows of the iteration's own E-step (one flipped row moves two centres by
            # |x - c| / n_k ~ 3e-3 here), so the band for labels_ is ~1e-3 of the distance; the kernel's own E-step
            # parity at fixed centres (5e-5 band) is test_every_lloyd_kernel_through_the_abi_selector.
            _labels_equal_mod_ties(tag, km.labels_, ref.labels_, X, ref.cluster_centers_, tie=2e-3, max_frac=1e-3)
            assert km.n_iter_ == ref.n_iter_
            _centres_close_mod_flips(tag + "_centers", km, ref, X, init=cur)
            record(tag + "_inertia", km.inertia_, ref.inertia_, 1e-5, 0)
            cur = ref.cluster_centers_.copy()


def test_float64_tensor_core_lloyd_pass_matches_sklearn_float64():
    """The float64 form of the tcgen05 pass (DIC_KM_KERNEL(6): tensor-core screen + exact float64 rows, float64 sums) on a
    reference-set-like matrix (uniform float64 draws, the operand of the gap statistic's fits) against scikit-learn's
    float64 Lloyd iteration from the same centres: labels identical up to float64 near-ties, centres to 1e-12."""
    import warnings
    from sklearn.cluster import KMeans
    from deep_interpolation_clustering_b200 import kmeans as km_mod
    orig = km_mod._Device.lloyd_run

    def forced(self, centers, flags, n_steps, tol):
        return orig(self, centers, flags | 6 << 8, n_steps, tol)
    rng = np.random.RandomState(11)
    X = rng.uniform(-1.0, 3.0, size=(200_003, 64))
    for K in (3, 10, 16):
        cur = X[:K].copy()
        for it in (1, 2):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                ref = KMeans(n_clusters=K, init=cur, n_init=1, max_iter=1, tol=0.0).fit(X)
            km_mod._Device.lloyd_run = forced
            try:
                km = km_mod.KMeansB200(n_clusters=K, init=cur, n_init=1, max_iter=1, tol=0.0).fit(X)
            finally:
                km_mod._Device.lloyd_run = orig
            tag = f"tc64_fit_K{K}_it{it}"
            _labels_equal_mod_ties(tag, km.labels_, ref.labels_, X, ref.cluster_centers_, tie=1e-9, max_frac=1e-4)
            assert km.n_iter_ == ref.n_iter_
            if np.array_equal(km.labels_, ref.labels_):
                np.testing.assert_allclose(km.cluster_centers_, ref.cluster_centers_, rtol=1e-12, atol=1e-12)
            record(tag + "_inertia", km.inertia_, ref.inertia_, 1e-10, 0)
            cur = ref.cluster_centers_.copy()


def test_config4_size_lloyd_matches_sklearn_with_fixed_init():
    """BASELINE config 4 size (1M x 64, K = 10): three Lloyd iterations from the same initial centres against
    scikit-learn run here on the host (sklearn/cluster/_k_means_lloyd.pyx:196-213 decides ties): labels equal except
    documented near-ties, inertia to 1e-5, centres to 1e-5 plus the displacement the near-tie rows themselves cause
    (_centres_close_mod_flips: 189 rows assigned differently move a centre of ~10^5 rows by up to ~4e-4).

    Near-tie bound at this size: sklearn decides a float32 row by ||c||^2 - 2 x.c evaluated in float32 (a 64-term dot
    product of magnitude ~10^3, i.e. ~1e-4 absolute noise on squared distances of ~60-100), so rows whose float64
    margin between the two nearest centres is below 2e-4 of the distance are decided by sklearn's own rounding (and
    by its thread count).  Measured here: 189 of 1,000,000 rows differ, worst margin 9.1e-5; none above the bound."""
    import warnings
    from sklearn.cluster import KMeans
    from deep_interpolation_clustering_b200 import synth
    from deep_interpolation_clustering_b200.kmeans import KMeansB200
    X = synth.make_blobs(1_000_000, 64, 5, seed=4)
    init = X[:10].copy()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = KMeans(n_clusters=10, init=init, n_init=1, max_iter=3, tol=0.0).fit(X)
    km = KMeansB200(n_clusters=10, init=init, n_init=1, max_iter=3, tol=0.0).fit(X)
    _labels_equal_mod_ties("c4_1M_labels", km.labels_, ref.labels_, X, ref.cluster_centers_, tie=2e-4, max_frac=1e-3)
    assert km.n_iter_ == ref.n_iter_
    _centres_close_mod_flips("c4_1M_centers", km, ref, X)
    record("c4_1M_inertia", km.inertia_, ref.inertia_, 1e-5, 0)
