#!/usr/bin/env python
"""Secondary benchmark: the K-selection path (BASELINE config 4): Lloyd iterations, k-means++
seeding, the pairwise-distance "inertia" and a full gap-statistic sweep on one B200.

    python benchmarks/bench_kselect.py [--n 1000000] [--d 64] [--sweep-n 100000]

Prints one JSON object.  CPU figures (scikit-learn on the host cores) are taken at sizes where
the reference can run at all: its pairwise step materialises an n_c x n_c matrix.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_interpolation_clustering_b200 import synth  # noqa: E402
from deep_interpolation_clustering_b200.gap import KM, pairwise_dist_sum  # noqa: E402
from deep_interpolation_clustering_b200.kmeans import KMeansB200, _Device  # noqa: E402


def timed(fn, reps=3):
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--d", type=int, default=64)
    ap.add_argument("--sweep-n", type=int, default=100_000)
    ap.add_argument("--cpu-n", type=int, default=4000)
    ap.add_argument("--full-refs", type=int, default=0, help="reference sets of the full-size (c4) sweep; 0 = skip")
    ap.add_argument("--full-ninit", type=int, default=2)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    out = {"n": args.n, "d": args.d}
    X = torch.from_numpy(synth.make_blobs(args.n, args.d, 5, seed=4)).to(dev)

    # one Lloyd pass (kernel only), K = 4 and 10
    for K in (4, 10):
        st = _Device(X, K)
        cen = X[:K].clone().contiguous()
        st.assign(cen, 0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            st.assign(cen, 1)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        out[f"lloyd_pass_ms_K{K}"] = round(ms, 4)
        out[f"lloyd_pass_GBps_K{K}"] = round(args.n * args.d * 4 / 1e9 / (ms * 1e-3), 1)

    for K in (4, 10):
        t, km = timed(lambda: KMeansB200(n_clusters=K, n_init=1, random_state=0).fit(X), reps=2)
        out[f"fit_s_K{K}"] = round(t, 4)
        out[f"fit_iters_K{K}"] = km.n_iter_
    t, km = timed(lambda: KMeansB200(n_clusters=4, n_init=10, random_state=0).fit(X), reps=1)
    out["fit_s_K4_ninit10"] = round(t, 4)

    # pairwise inertia kernel
    for n in (16384, 131072, 500000):
        if n > args.n:
            continue
        Xc = X[:n].contiguous()
        t, s = timed(lambda: pairwise_dist_sum(Xc), reps=2)
        out[f"pairwise_s_n{n}"] = round(t, 4)
        out[f"pairwise_Gdist_per_s_n{n}"] = round(n * n / 2 / t / 1e9, 1)

    # full sweep on device
    Xs = X[:args.sweep_n].contiguous()
    np.random.seed(123)
    km = KM(10, None, [], 10, 20)
    t0 = time.perf_counter()
    df = km.compute_gap_internal_metric(KMeansB200(n_init=10), Xs, k_max=10, n_references=20, version=1)
    torch.cuda.synchronize()
    out["gap_sweep_s"] = round(time.perf_counter() - t0, 2)
    out["gap_sweep_n"] = args.sweep_n
    out["gap_best_k"] = int(df["gap"].astype(float).idxmax())

    # BASELINE config 4 at full size (N x d, K = 2..10) with a reduced number of reference sets / restarts;
    # the full 20 x 10 sweep is (21 / (refs + 1)) * (10 / n_init) times this
    if args.full_refs > 0:
        km = KM(10, None, [], args.full_ninit, args.full_refs)
        t0 = time.perf_counter()
        df = km.compute_gap_internal_metric(KMeansB200(n_init=args.full_ninit), X, k_max=10,
                                            n_references=args.full_refs, version=1, draw="device")
        torch.cuda.synchronize()
        out["c4_sweep_s"] = round(time.perf_counter() - t0, 2)
        out["c4_sweep_cfg"] = {"n": args.n, "d": args.d, "k": "2..10", "n_references": args.full_refs,
                               "n_init": args.full_ninit, "draws": "device"}
        out["c4_best_k"] = int(df["gap"].astype(float).idxmax())

    # CPU: the reference's own stack at a size it can run
    from sklearn.cluster import KMeans
    from oracle import kmeans_oracle
    Xh = X[:args.cpu_n].cpu().numpy()
    np.random.seed(123)
    t0 = time.perf_counter()
    kmeans_oracle.gap_statistic(lambda k, data: KMeans(n_clusters=k, n_init=10).fit_predict(data), Xh, k_max=10,
                                n_references=20, version=1)
    out["cpu_gap_sweep_s"] = round(time.perf_counter() - t0, 2)
    out["cpu_gap_sweep_n"] = args.cpu_n
    out["cpu_cores"] = os.cpu_count()
    Xh = X.cpu().numpy()
    t0 = time.perf_counter()
    ref = KMeans(n_clusters=10, n_init=1, random_state=0).fit(Xh)
    dt = time.perf_counter() - t0
    out["cpu_sklearn_fit_s_K10"] = round(dt, 3)
    out["cpu_sklearn_ms_per_iter_K10"] = round(dt / ref.n_iter_ * 1e3, 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
