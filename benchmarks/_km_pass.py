"""Times one Lloyd pass (dic_kmeans_assign with the hot-loop flags) per kernel (DIC_KM_KERNEL selector), K, D and dtype."""
import sys, json, torch
sys.path.insert(0, '.')
from deep_interpolation_clustering_b200 import synth
from deep_interpolation_clustering_b200.kmeans import _Device
KERNELS = (("auto", 0), ("tile2", 1), ("rw", 2), ("tile", 3), ("tc", 5), ("tc64", 6))
out = {}
for D, N in ((64, 1_000_000), (128, 500_000), (256, 500_000)):
    X32 = torch.from_numpy(synth.make_blobs(N, D, 5, seed=4)).cuda()
    for name, X in (("f32", X32), ("f64", X32.double())):
        if name == "f64" and D > 64:
            continue
        for K in (2, 4, 8, 10, 16):
            for kern, sel in KERNELS:
                st = _Device(X, K)
                cen = X[:K].clone().contiguous()
                flags = 1 | 4 | sel << 8
                try:
                    st.assign(cen, flags)
                except ValueError:            # the kernel does not cover this shape
                    continue
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(20):
                    st.assign(cen, flags)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 20
                gb = N * D * X.element_size() / 1e9
                out[f"{name}_N{N}_D{D}_K{K}_{kern}"] = {"ms": round(ms, 4), "GBps": round(gb / (ms * 1e-3), 0)}
print(json.dumps(out, indent=0))
