"""Time per Lloyd iteration inside KMeansB200.fit (batched run: assign + partial-sum finish + centre update per iteration)
against the time of the bare pass: what the launches around the pass cost."""
import sys, json, time, torch, numpy as np
sys.path.insert(0, '.')
from deep_interpolation_clustering_b200.kmeans import KMeansB200
out = {}
for dt in (np.float64, np.float32):
    rng = np.random.RandomState(0)
    X = torch.from_numpy(rng.uniform(size=(1_000_000, 64)).astype(dt)).cuda()
    for K in (4, 10):
        init = X[:K].cpu().numpy().copy()
        for iters in (20, 120):
            km = KMeansB200(n_clusters=K, init=init, n_init=1, max_iter=iters, tol=0.0)
            km.fit(X)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            km.fit(X)
            torch.cuda.synchronize()
            out[f"{np.dtype(dt).name}_K{K}_it{iters}"] = {"s": round(time.perf_counter() - t0, 5), "n_iter": int(km.n_iter_)}
        a, b = out[f"{np.dtype(dt).name}_K{K}_it20"], out[f"{np.dtype(dt).name}_K{K}_it120"]
        if b["n_iter"] > a["n_iter"]:
            out[f"{np.dtype(dt).name}_K{K}_ms_per_iteration"] = round((b["s"] - a["s"]) / (b["n_iter"] - a["n_iter"]) * 1e3, 4)
print(json.dumps(out))
