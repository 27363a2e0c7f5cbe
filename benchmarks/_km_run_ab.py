import json, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from deep_interpolation_clustering_b200 import synth
from deep_interpolation_clustering_b200.kmeans import _Device
out = {}
N, D, K = 1_000_000, 64, 4
X = torch.from_numpy(synth.make_blobs(N, D, 4, seed=9)).cuda()
for name, sel in (("tc", 5), ("tile2", 1)):
    st = _Device(X, K)
    for tag, tol in (("converges_early", 1e-4), ("never_converges", 0.0)):
        res = []
        for rep in range(3):
            cen = X[:K].clone().contiguous()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            r = st.lloyd_run(cen, 1 | 4 | sel << 8, 8, tol)
            e1.record()
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            res.append((round((t1 - t0) * 1e3, 3), round(e0.elapsed_time(e1), 3)))
        out[f"{name}_{tag}"] = {"host_ms_gpu_ms": res, "ret": str(r)[:80]}
print(json.dumps(out))
