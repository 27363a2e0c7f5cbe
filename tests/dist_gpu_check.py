"""Multi-GPU (NCCL) parity of the sharded hot path against the single-GPU result.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/dist_gpu_check.py

Every rank computes the FULL problem on its own GPU (the single-process truth) and its SHARD through the
sharded entry points; the shard results must reproduce the matching rows / the global
reductions: interpolation parameter gradients (one packed all-reduce), DEC target distribution and fused KL step
(global column sum, global 'batchmean'), k-means (labels bit-exact, centres, inertia).  Prints one JSON line on
rank 0 and exits non-zero on any mismatch.  Run by tests/test_gpu_multi.py when >= 2 GPUs are visible.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from deep_interpolation_clustering_b200 import dist_check                          # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    report = dist_check.run(rank, world, dev)
    dist.barrier()
    if rank == 0:
        print(json.dumps({"world": world, "ok": True, **{k: float(f"{v:.3e}") for k, v in report.items()}}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
