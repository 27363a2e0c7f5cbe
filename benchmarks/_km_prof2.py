"""ncu target: a few Lloyd passes (hot-loop flags) at 1M x 64, K = 4 and 10."""
import sys, torch
sys.path.insert(0, '.')
from deep_interpolation_clustering_b200 import synth
from deep_interpolation_clustering_b200.kmeans import _Device
X = torch.from_numpy(synth.make_blobs(1_000_000, 64, 5, seed=4)).cuda()
for K in (4, 10):
    st = _Device(X, K)
    cen = X[:K].clone().contiguous()
    for _ in range(3):
        st.assign(cen, 5)
torch.cuda.synchronize()
