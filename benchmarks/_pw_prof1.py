import sys, torch
sys.path.insert(0, '.')
from deep_interpolation_clustering_b200 import synth
from deep_interpolation_clustering_b200.gap import pairwise_dist_sum
X = torch.from_numpy(synth.make_blobs(65536, 64, 5, seed=4)).cuda()
for _ in range(3):
    s = pairwise_dist_sum(X)
torch.cuda.synchronize()
print(float(s))
