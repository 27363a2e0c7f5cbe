import sys, torch
sys.path.insert(0, '.')
from deep_interpolation_clustering_b200 import synth
from deep_interpolation_clustering_b200.kmeans import _Device
D, N, K = 64, 1_000_000, 10
X = torch.from_numpy(synth.make_blobs(N, D, 5, seed=4)).cuda()
st = _Device(X, K)
cen = X[:K].clone().contiguous()
for _ in range(3):
    st.assign(cen, 1 | 4 | 5 << 8)
torch.cuda.synchronize()
