"""c4 K-selection sweep with its (k, reference set) fits dealt to the GPUs of one node.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 \
        benchmarks/bench_gap_tasks.py [--n 1000000 --d 64 --refs 4 --ninit 2]

Every rank holds the whole (n, d) matrix; the only exchange is one all-reduce of the inertia table.  Prints one JSON
line on rank 0: wall time of the sweep (barrier on both sides, max over ranks) and the K the gap statistic picks.
Also runs single-process (N = 1, plain `python`).
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from deep_interpolation_clustering_b200 import synth                 # noqa: E402
from deep_interpolation_clustering_b200.gap import KM                # noqa: E402
from deep_interpolation_clustering_b200.kmeans import KMeansB200     # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--d", type=int, default=64)
    ap.add_argument("--refs", type=int, default=4)
    ap.add_argument("--ninit", type=int, default=2)
    ap.add_argument("--kmax", type=int, default=10)
    ap.add_argument("--draw", default="device", choices=["device", "device32"])
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    X = torch.from_numpy(synth.make_blobs(args.n, args.d, 5, seed=3).astype(np.float32)).to(dev)
    km = KM(args.kmax, None, [], args.ninit, args.refs)

    def sweep():
        return km.compute_gap_internal_metric(KMeansB200(n_init=args.ninit, random_state=1, device=dev), X,
                                              k_max=args.kmax, n_references=args.refs, version=1, draw=args.draw,
                                              task_parallel=True, group=dist.group.WORLD if world > 1 else None)
    km_small = KM(3, None, [], 1, 1)                                  # warm-up: allocator, NCCL communicator
    km_small.compute_gap_internal_metric(KMeansB200(n_init=1, random_state=1, device=dev), X[:50000], k_max=3,
                                         n_references=1, version=1, draw="device", task_parallel=True,
                                         group=dist.group.WORLD if world > 1 else None)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    df = sweep()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        best = int(df["k"][df["gap"].astype(float).idxmax()])
        print(json.dumps({"workload": "c4 gap sweep, tasks dealt to the GPUs", "n": args.n, "d": args.d,
                          "k": f"2..{args.kmax}", "n_references": args.refs, "n_init": args.ninit, "draw": args.draw, "n_gpus": world,
                          "sweep_s": round(float(dt), 3), "best_k": best,
                          "gap": [round(float(g), 5) for g in df["gap"]]}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
