"""Multi-GPU plumbing: one process per GPU, encounters sharded by contiguous index blocks,
small NCCL all-reduces only where the path has a real exchange step.

    interpolation fwd/bwd     no data-path collective; parameter gradients (C + C^2 + C floats)
                              ride in ONE packed all-reduce          -> allreduce_gradients
    DEC target distribution   f_j = sum_i q_ij over the GLOBAL batch -> sharded_target_distribution
    DEC KL step               all-reduce(K doubles) between the two kernels, then the centre
                              gradients (K*D floats) and the KL sum  -> sharded_dec_kl_step
    k-means Lloyd             per-iteration all-reduce of the packed [sums | counts | stats]
                              buffer                                 -> KMeansB200(process_group=...)

The reference only knows single-process nn.DataParallel (pretrain_trainer.py:21), under which
target_distribution would be normalised per replica chunk; the functions here define the
global-batch semantic explicitly.  The collectives are torch.distributed calls (NCCL on GPUs,
gloo in the CPU tests), so the host logic is testable without a GPU by injecting the compute.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

__all__ = ["shard_range", "allreduce_gradients", "sharded_target_distribution", "sharded_dec_kl_step",
           "world"]


def world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_range(n, rank, world_size):
    """Contiguous block [lo, hi) of n units owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(n), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_gradients(params, group=None, average=False):
    """Sum (or average) the .grad of `params` across ranks with ONE packed all-reduce.
    The hot path's own parameters are 6 + 36 + 6 floats, so the message is latency bound."""
    params = [p for p in params if p.grad is not None]
    if not params:
        return
    rank, ws = world(group)
    if ws == 1:
        return
    flat = torch.cat([p.grad.reshape(-1).to(torch.float32) for p in params])
    dist.all_reduce(flat, group=group)
    if average:
        flat /= ws
    off = 0
    for p in params:
        n = p.grad.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n


def sharded_target_distribution(q, group=None, colsum_fn=None, p_fn=None):
    """target_distribution (dec.py:66-76) for a batch whose rows are sharded across ranks:
    p is normalised with the all-reduced column sum, so every rank gets exactly the rows of the
    single-process result on the concatenated batch."""
    if colsum_fn is None or p_fn is None:
        from . import functional as F_
        colsum_fn = colsum_fn or F_.colsum
        p_fn = p_fn or F_.dec_target_distribution
    f = colsum_fn(q)
    if world(group)[1] > 1:
        dist.all_reduce(f, group=group)
    return p_fn(q, f)


def sharded_dec_kl_step(z, mu, alpha=1.0, weight=1.0, group=None, assign_fn=None, kl_fn=None):
    """Fused DEC step on a row shard: q and the local column sum, all-reduce(K), then p / KL /
    closed-form gradients with the GLOBAL batch size in the 'batchmean' divisor; the centre
    gradient and the KL value are all-reduced in one packed message.  grad_z stays local."""
    if assign_fn is None or kl_fn is None:
        from . import functional as F_
        assign_fn = assign_fn or F_.dec_assign
        kl_fn = kl_fn or F_.dec_kl_from_colsum
    rank, ws = world(group)
    out = assign_fn(z, mu, alpha)
    # one packed message: [column sum (K) | local batch size]
    head = torch.cat([out["colsum"].to(torch.float64), torch.tensor([z.shape[0]], dtype=torch.float64,
                                                                    device=out["colsum"].device)])
    if ws > 1:
        dist.all_reduce(head, group=group)
    f, n_global = head[:-1].contiguous(), int(round(float(head[-1])))
    out.update(kl_fn(z, mu, f, alpha, weight=weight, batch=n_global))
    packed = torch.cat([out["grad_mu"].reshape(-1).to(torch.float64), out["kl"].reshape(-1).to(torch.float64)])
    if ws > 1:
        dist.all_reduce(packed, group=group)
    out["grad_mu"] = packed[:-1].view_as(out["grad_mu"]).to(out["grad_mu"].dtype)
    out["kl"] = packed[-1:]
    out["colsum"] = f
    out["batch"] = n_global
    return out
