"""Kernel-time table of a whole fit (torch profiler): argv = dtype K n_init  (uniform 1M x 64 rows: a reference set)."""
import json, sys, time
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, '.')
from deep_interpolation_clustering_b200 import kmeans as km_mod
dt = np.float64 if (len(sys.argv) < 2 or sys.argv[1] == "f64") else np.float32
K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
ninit = int(sys.argv[3]) if len(sys.argv) > 3 else 2
X = torch.from_numpy(np.random.RandomState(0).uniform(size=(1_000_000, 64)).astype(dt)).cuda()
km_mod.KMeansB200(n_clusters=K, n_init=ninit, random_state=0).fit(X); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    t0 = time.perf_counter()
    km = km_mod.KMeansB200(n_clusters=K, n_init=ninit, random_state=0).fit(X); torch.cuda.synchronize()
    wall = time.perf_counter() - t0
rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0 and not e.key.startswith("aten::")]
rows.sort(key=lambda r: -r[2])
tot = sum(r[2] for r in rows)
print(f"wall {wall*1e3:.2f} ms, device time {tot/1e3:.2f} ms, n_iter {km.n_iter_}")
for k, c, t in rows[:14]:
    print(f"   {t/1e3:8.3f} ms  x{c:5d}  {k[:100]}")
