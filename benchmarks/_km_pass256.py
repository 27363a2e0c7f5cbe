"""Lloyd pass timing at the reference's real latent width (D = 256) and at D = 128, float32."""
import sys, json, torch
sys.path.insert(0, '.')
from deep_interpolation_clustering_b200 import synth
from deep_interpolation_clustering_b200.kmeans import _Device
out = {}
for D in (256, 128):
    X = torch.from_numpy(synth.make_blobs(500_000, D, 5, seed=4)).cuda()
    for K in (4, 10, 16):
        st = _Device(X, K)
        cen = X[:K].clone().contiguous()
        st.assign(cen, 0)
        for flags in (5,):
            st.assign(cen, flags)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                st.assign(cen, flags)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            out[f"D{D}_K{K}"] = {"ms": round(ms, 4), "GBps": round(X.numel() * 4 / ms / 1e6, 1)}
print(json.dumps(out))
