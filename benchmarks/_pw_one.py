"""Two calls of the pairwise-distance sum at n = 131072, D = 64 (ncu target: skip the first launch)."""
import sys, torch
sys.path.insert(0, '.')
from deep_interpolation_clustering_b200 import synth
from deep_interpolation_clustering_b200.gap import pairwise_dist_sum
X = torch.from_numpy(synth.make_blobs(131072, 64, 5, seed=4)).cuda()
for _ in range(2):
    s = pairwise_dist_sum(X)
torch.cuda.synchronize()
print(float(s))
