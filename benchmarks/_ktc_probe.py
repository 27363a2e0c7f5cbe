import sys, json, torch
sys.path.insert(0, '.')
from deep_interpolation_clustering_b200 import synth
from deep_interpolation_clustering_b200.kmeans import _Device
out = {}
for D, N, dt, sel in ((64, 1_000_000, torch.float64, 6),):
    X = torch.from_numpy(synth.make_blobs(N, D, 5, seed=4)).cuda().to(dt)
    for K in (2, 4, 10, 16):
        for ws in (True, False):
            st = _Device(X, K)
            cen = X[:K].clone().contiguous()
            flags = 1 | 4 | sel << 8
            st.assign(cen, flags, want_sums=ws)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                st.assign(cen, flags, want_sums=ws)
            e1.record()
            torch.cuda.synchronize()
            out[f"K{K}_sums{int(ws)}"] = round(e0.elapsed_time(e1) / 20, 4)
print(json.dumps(out))
