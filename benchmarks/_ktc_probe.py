import sys, json, torch
sys.path.insert(0, '.')
from deep_interpolation_clustering_b200 import synth
from deep_interpolation_clustering_b200.kmeans import _Device
out = {}
for D, N in ((64, 1_000_000), (128, 500_000), (128, 1_000_000), (256, 250_000), (256, 500_000)):
    X = torch.from_numpy(synth.make_blobs(N, D, 5, seed=4)).cuda()
    for K in (10,):
        for ws in (True, False):
            st = _Device(X, K)
            cen = X[:K].clone().contiguous()
            flags = 1 | 4 | 5 << 8
            st.assign(cen, flags, want_sums=ws)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                st.assign(cen, flags, want_sums=ws)
            e1.record()
            torch.cuda.synchronize()
            out[f"D{D}_N{N}_sums{int(ws)}"] = round(e0.elapsed_time(e1) / 20, 4)
print(json.dumps(out))
