"""Whole fits with the Lloyd iterations on the dispatched kernel against the same fits forced onto the CUDA-core tile kernel."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from deep_interpolation_clustering_b200 import synth
from deep_interpolation_clustering_b200 import kmeans as km_mod

orig = km_mod._Device.lloyd_run
def forced(sel):
    def f(self, centers, flags, n_steps, tol):
        return orig(self, centers, flags | sel << 8, n_steps, tol)
    return f

def timed(fn):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize()
    return time.perf_counter() - t0, out

out = {}
for N, D, K, ninit in ((500_000, 256, 4, 20), (500_000, 256, 10, 1), (1_000_000, 64, 4, 10), (1_000_000, 64, 10, 1)):
    X = torch.from_numpy(synth.make_blobs(N, D, 4, seed=9)).cuda()
    for name, sel in (("auto", 0), ("tile2", 1)):
        km_mod._Device.lloyd_run = forced(sel) if sel else orig
        t, km = timed(lambda: km_mod.KMeansB200(n_clusters=K, n_init=ninit, random_state=0).fit(X))
        out[f"N{N}_D{D}_K{K}_ninit{ninit}_{name}"] = {"s": round(t, 4), "iters": int(km.n_iter_)}
km_mod._Device.lloyd_run = orig
print(json.dumps(out))
