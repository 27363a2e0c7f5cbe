import json, sys, time
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, '.')
from deep_interpolation_clustering_b200 import synth
from deep_interpolation_clustering_b200 import kmeans as km_mod
orig = km_mod._Device.lloyd_run
def forced(sel):
    def f(self, centers, flags, n_steps, tol):
        return orig(self, centers, flags | sel << 8, n_steps, tol)
    return f
N, D, K = 1_000_000, 64, 4
X = torch.from_numpy(synth.make_blobs(N, D, 4, seed=9)).cuda()
for name, sel in (("auto", 0), ("tile2", 1)):
    km_mod._Device.lloyd_run = forced(sel) if sel else orig
    km_mod.KMeansB200(n_clusters=K, n_init=3, random_state=0).fit(X); torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        t0 = time.perf_counter()
        km_mod.KMeansB200(n_clusters=K, n_init=3, random_state=0).fit(X); torch.cuda.synchronize()
        wall = time.perf_counter() - t0
    print(name, "wall ms", round(wall * 1e3, 2))
    rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0]
    rows.sort(key=lambda r: -r[2])
    for k, c, t in rows[:12]:
        print(f"   {t/1e3:8.3f} ms  x{c:4d}  {k[:90]}")
