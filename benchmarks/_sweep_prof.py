import sys, time, cProfile, pstats, torch, numpy as np
sys.path.insert(0, '.')
from deep_interpolation_clustering_b200 import synth
from deep_interpolation_clustering_b200.gap import KM
from deep_interpolation_clustering_b200.kmeans import KMeansB200
X = torch.from_numpy(synth.make_blobs(20000, 64, 5, seed=4)).cuda()
np.random.seed(123)
km = KM(10, None, [], 10, 20)
km.compute_gap_internal_metric(KMeansB200(n_init=2), X, k_max=4, n_references=2, version=1)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable(); t0 = time.perf_counter()
df = km.compute_gap_internal_metric(KMeansB200(n_init=10), X, k_max=10, n_references=5, version=1)
torch.cuda.synchronize(); dt = time.perf_counter() - t0; pr.disable()
print('sweep (5 refs)', dt)
pstats.Stats(pr).sort_stats('tottime').print_stats(25)
