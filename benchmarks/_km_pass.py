"""Times one Lloyd pass (dic_kmeans_assign with the hot-loop flags) for the two kernels, K and dtype."""
import sys, json, torch
sys.path.insert(0, '.')
from deep_interpolation_clustering_b200 import synth
from deep_interpolation_clustering_b200.kmeans import _Device
X32 = torch.from_numpy(synth.make_blobs(1_000_000, 64, 5, seed=4)).cuda()
out = {}
for name, X in (("f32", X32), ("f64", X32.double())):
    for K in (2, 4, 8, 10, 16):
        for kern, sel in (("tile2", 1), ("rw", 2), ("tile", 3)):      # DIC_KM_KERNEL(sel), include/dic_b200.h
            st = _Device(X, K)
            cen = X[:K].clone().contiguous()
            try:
                st.assign(cen, sel << 8)
            except ValueError:            # the kernel does not cover this shape
                continue
            for flags in (5 | sel << 8, 1 | sel << 8):
                st.assign(cen, flags)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(20):
                    st.assign(cen, flags)
                e1.record()
                torch.cuda.synchronize()
                out[f"{name}_K{K}_{kern}_flags{flags & 255}"] = round(e0.elapsed_time(e1) / 20, 4)
print(json.dumps(out, indent=0))
