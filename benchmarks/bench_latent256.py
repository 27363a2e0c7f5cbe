"""The K-selection / initialisation path at the reference's real latent width (D = 256, pretrain_interp.py:97-100):
p3's KMeans(n_clusters=4, n_init=20) initialisation (clustering_trainer.py:75-82), a K = 10 fit, the pairwise
"inertia" and the silhouette.  Prints one JSON line."""
import json, sys, time
import numpy as np
import torch
sys.path.insert(0, '.')
from deep_interpolation_clustering_b200 import synth
from deep_interpolation_clustering_b200.gap import KM
from deep_interpolation_clustering_b200.internal_eval import Sihouette
from deep_interpolation_clustering_b200.kmeans import KMeansB200

def timed(fn, reps=1):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, out

N, D = 500_000, 256
X = torch.from_numpy(synth.make_blobs(N, D, 4, seed=9)).cuda()
out = {"n": N, "d": D}
t, km = timed(lambda: KMeansB200(n_clusters=4, n_init=20, random_state=0).fit(X))
out["fit_K4_ninit20_s"] = round(t, 4); out["fit_K4_iters"] = int(km.n_iter_)
t, km10 = timed(lambda: KMeansB200(n_clusters=10, n_init=1, random_state=0).fit(X))
out["fit_K10_s"] = round(t, 4); out["fit_K10_iters"] = int(km10.n_iter_)
labels = km.labels_
t, w = timed(lambda: KM(4).compute_inertia_v1(labels, X))
out["inertia_v1_s"] = round(t, 4); out["inertia_v1"] = w
Xs, ls = X[:100_000], labels[:100_000]
t, sil = timed(lambda: Sihouette()(Xs, ls.cpu().numpy()))
out["silhouette_n100k_s"] = round(t, 4); out["silhouette"] = round(sil, 6)
print(json.dumps(out))
