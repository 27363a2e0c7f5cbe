"""GPU parity of the cluster-validity metrics the gap loop evaluates for every K: the four classes of
internal_eval.py:15-147 on the device kernels (dic_cluster_scatter, dic_dunn_minmax, dic_cluster_rowsums) against
  * the values the reference's own classes produced (tests/golden/internal_eval.npz, oracle/gen_golden.py), and
  * scikit-learn / the numpy oracle at sizes the reference's pure-Python Dunn loop cannot reach."""
import numpy as np
import pytest
import torch

from gpu_util import record

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", ["blobs", "dups"])
@pytest.mark.parametrize("dtag", ["f32", "f64"])
def test_metrics_match_the_reference_classes(golden, tag, dtag):
    from deep_interpolation_clustering_b200 import internal_eval as ie
    g = golden("internal_eval")
    X = g[tag + "_X"].astype(np.float32 if dtag == "f32" else np.float64)
    lab = g[tag + "_labels"]
    pre = f"ieval_{tag}_{dtag}_"
    record(pre + "dunn", ie.DunnIndex()(X, lab), float(g[tag + "_dunn"]), 1e-5, 0)
    record(pre + "ch", ie.CHIndex()(X, lab), float(g[tag + "_ch"]), 1e-5, 0)
    record(pre + "db", ie.DBIndex()(X, lab), float(g[tag + "_db"]), 1e-5, 0)
    record(pre + "silhouette", ie.Sihouette()(X, lab), float(g[tag + "_silhouette"]), 1e-5, 1e-6)


def test_dunn_drops_touching_cluster_pairs_like_nonzero(golden):
    """internal_eval.py:106: a pair of clusters whose nearest distance is exactly 0 drops out of the minimum; when
    every pair touches the reference raises ValueError (min of an empty sequence) and so does the mirror."""
    from deep_interpolation_clustering_b200 import internal_eval as ie
    X = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 0.0], [3.0, 1.0]], np.float32)
    with pytest.raises(ValueError):
        ie.DunnIndex()(X, np.array([0, 0, 1, 1]))


@pytest.mark.parametrize("N,D,K", [(200_000, 64, 10), (50_000, 256, 4), (3001, 20, 7)])
def test_ch_db_match_sklearn_at_size(N, D, K):
    from sklearn import metrics
    from deep_interpolation_clustering_b200 import internal_eval as ie, synth
    X = synth.make_blobs(N, D, 5, seed=N % 97)
    lab = np.random.RandomState(1).randint(0, K, size=N)
    lab[:K] = np.arange(K)
    Xd = torch.from_numpy(X).cuda()
    record(f"ch_N{N}_D{D}", ie.CHIndex()(Xd, lab), metrics.calinski_harabasz_score(X.astype(np.float64), lab), 1e-5, 0)
    record(f"db_N{N}_D{D}", ie.DBIndex()(Xd, lab), metrics.davies_bouldin_score(X.astype(np.float64), lab), 1e-5, 0)


@pytest.mark.parametrize("N,D,K,dtag", [(3000, 64, 6, "f32"), (2500, 33, 3, "f64"), (130, 8, 32, "f32")])
def test_dunn_matches_the_oracle_at_size(N, D, K, dtag):
    from oracle import kmeans_oracle
    from deep_interpolation_clustering_b200 import internal_eval as ie, synth
    X = synth.make_blobs(N, D, 4, seed=D).astype(np.float32 if dtag == "f32" else np.float64)
    lab = np.random.RandomState(2).randint(0, K, size=N)
    lab[:K] = np.arange(K)
    want = kmeans_oracle.dunn_index(X.astype(np.float64), lab)
    record(f"dunn_N{N}_D{D}_{dtag}", ie.DunnIndex()(torch.from_numpy(X).cuda(), lab), want, 1e-5, 0)


def test_metrics_inside_the_gap_dataframe_cover_dunn(golden):
    """KM(..., internal_metrics incl. Dunn_Index) fills every metric column (p2_clustering_optK.py:401-405)."""
    from deep_interpolation_clustering_b200.gap import KM
    from deep_interpolation_clustering_b200.kmeans import KMeansB200
    g = golden("gap")
    names = ["Dunn_Index", "Sihouette", "Davies-Bouldin_Index", "Calinski-Harabasz"]
    km = KM(4, None, names, 1, 2)
    df = km.compute_gap_internal_metric(KMeansB200(n_init=1, random_state=0), g["X"], k_max=4, n_references=2, version=1)
    assert list(df.columns[5:]) == names and np.isfinite(df[names].to_numpy(dtype=np.float64)).all()
