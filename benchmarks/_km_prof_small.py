import sys, time, cProfile, pstats, torch
sys.path.insert(0, '.')
from deep_interpolation_clustering_b200 import synth
from deep_interpolation_clustering_b200.kmeans import KMeansB200
dev = torch.device('cuda:0')
X = torch.from_numpy(synth.make_blobs(20000, 64, 5, seed=4)).to(dev)
KMeansB200(n_clusters=6, n_init=10, random_state=0).fit(X)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
t0 = time.perf_counter()
for k in (4, 8):
    km = KMeansB200(n_clusters=k, n_init=10, random_state=0).fit(X.double())
torch.cuda.synchronize()
dt = time.perf_counter() - t0
pr.disable()
print('2 fits (n_init=10 each)', dt, 'iters', km.n_iter_)
pstats.Stats(pr).sort_stats('tottime').print_stats(30)
