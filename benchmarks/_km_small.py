"""Lloyd pass time against N for the CUDA-core tile kernel and the tcgen05 kernel (dispatch threshold)."""
import sys, json, torch
sys.path.insert(0, '.')
from deep_interpolation_clustering_b200 import synth
from deep_interpolation_clustering_b200.kmeans import _Device
out = {}
import os
F64 = os.environ.get("KM_SMALL_F64") == "1"
for D in ((64,) if F64 else (64, 128, 256)):
    for N in (10_000, 20_000, 50_000, 100_000, 300_000):
        X = torch.from_numpy(synth.make_blobs(N, D, 5, seed=4)).cuda()
        if F64:
            X = X.double()
        for K in (2, 4, 10, 16):
            for kern, sel in (("tile2", 1), ("tc", 6 if F64 else 5)):
                st = _Device(X, K)
                cen = X[:K].clone().contiguous()
                flags = 1 | 4 | sel << 8
                st.assign(cen, flags)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for _ in range(20):
                        st.assign(cen, flags)
                g.replay()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                g.replay()
                e1.record()
                torch.cuda.synchronize()
                out[f"D{D}_N{N}_K{K}_{kern}"] = round(e0.elapsed_time(e1) / 20 * 1e3, 1)
print(json.dumps(out))
