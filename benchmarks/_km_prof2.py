import sys, torch
sys.path.insert(0, '.')
from deep_interpolation_clustering_b200 import synth
from deep_interpolation_clustering_b200.kmeans import _Device
X = torch.from_numpy(synth.make_blobs(1_000_000, 64, 5, seed=4)).cuda()
st = _Device(X, 10)
cen = X[:10].clone().contiguous().cpu().numpy()
for _ in range(3):
    st.assign_hc(cen, 5)
torch.cuda.synchronize()
