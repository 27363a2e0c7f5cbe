"""One Lloyd pass of the dispatched kernel at 1M x 64, K = 10 (ncu target).   python benchmarks/_km_one.py [f32|f64] [K]"""
import sys, torch
sys.path.insert(0, '.')
from deep_interpolation_clustering_b200 import synth
from deep_interpolation_clustering_b200.kmeans import _Device
dt = sys.argv[1] if len(sys.argv) > 1 else "f32"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
X = torch.from_numpy(synth.make_blobs(1_000_000, 64, 5, seed=4)).cuda()
if dt == "f64":
    X = X.double()
st = _Device(X, K)
cen = X[:K].clone().contiguous()
for _ in range(3):
    st.assign(cen, 1 | 4)
torch.cuda.synchronize()
