"""Time per Lloyd iteration inside KMeansB200.fit at a launch-bound size (20,000 x 64)."""
import sys, json, time, torch, numpy as np
sys.path.insert(0, '.')
from deep_interpolation_clustering_b200.kmeans import KMeansB200
out = {}
for dt in (np.float64, np.float32):
    X = torch.from_numpy(np.random.RandomState(0).uniform(size=(20_000, 64)).astype(dt)).cuda()
    for K in (4, 10):
        init = X[:K].cpu().numpy().copy()
        r = {}
        for iters in (40, 240):
            km = KMeansB200(n_clusters=K, init=init, n_init=1, max_iter=iters, tol=0.0)
            km.fit(X); torch.cuda.synchronize()
            t0 = time.perf_counter(); km.fit(X); torch.cuda.synchronize()
            r[iters] = (time.perf_counter() - t0, int(km.n_iter_))
        if r[240][1] > r[40][1]:
            out[f"{np.dtype(dt).name}_K{K}_us_per_iteration"] = round((r[240][0] - r[40][0]) / (r[240][1] - r[40][1]) * 1e6, 2)
        out[f"{np.dtype(dt).name}_K{K}_n_iter"] = r[240][1]
print(json.dumps(out))
