"""Times one Lloyd pass (dic_kmeans_assign with the hot-loop flags) for the two kernels, K and dtype."""
import os, sys, json, torch
sys.path.insert(0, '.')
from deep_interpolation_clustering_b200 import synth
from deep_interpolation_clustering_b200.kmeans import _Device
X32 = torch.from_numpy(synth.make_blobs(1_000_000, 64, 5, seed=4)).cuda()
out = {}
for name, X in (("f32", X32), ("f64", X32.double())):
    for K in (2, 4, 8, 10, 16):
        for kern in ("tile2", "rw", "tile"):
            os.environ.pop("DIC_KMEANS_NO_RW", None)
            os.environ.pop("DIC_KMEANS_NO_TILE2", None)
            if kern != "tile2":
                os.environ["DIC_KMEANS_NO_TILE2"] = "1"
            if kern == "tile":
                os.environ["DIC_KMEANS_NO_RW"] = "1"
            st = _Device(X, K)
            cen = X[:K].clone().contiguous()
            st.assign(cen, 0)
            for flags in (5, 1):
                st.assign(cen, flags)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(20):
                    st.assign(cen, flags)
                e1.record()
                torch.cuda.synchronize()
                out[f"{name}_K{K}_{kern}_flags{flags}"] = round(e0.elapsed_time(e1) / 20, 4)
print(json.dumps(out, indent=0))
