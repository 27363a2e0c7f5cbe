"""Single-launch ncu target for the tcgen05 Lloyd passes: argv[1] = f32 | f64 (1M x 64, K = 10, loop form of the pass)."""
import sys, torch
sys.path.insert(0, '.')
from deep_interpolation_clustering_b200 import synth
from deep_interpolation_clustering_b200.kmeans import _Device
f64 = len(sys.argv) > 1 and sys.argv[1] == "f64"
D, N, K = 64, 1_000_000, 10
X = torch.from_numpy(synth.make_blobs(N, D, 5, seed=4)).cuda()
if f64:
    X = X.double()
st = _Device(X, K)
cen = X[:K].clone().contiguous()
for _ in range(3):
    st.assign(cen, 1 | 4 | (6 if f64 else 5) << 8)
torch.cuda.synchronize()
