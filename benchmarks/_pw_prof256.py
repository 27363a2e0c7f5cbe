import sys, time, torch
sys.path.insert(0, '.')
from deep_interpolation_clustering_b200 import synth
from deep_interpolation_clustering_b200.gap import pairwise_dist_sum
for n, D in ((65536, 256), (65536, 128)):
    X = torch.from_numpy(synth.make_blobs(n, D, 5, seed=4)).cuda()
    for _ in range(2):
        s = pairwise_dist_sum(X)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        s = pairwise_dist_sum(X)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(n, D, float(s), f"{dt*1e3:.3f} ms", f"{n*n/2/dt/1e9:.1f} Gdist/s", flush=True)
