"""Multi-GPU (NCCL) parity of the sharded hot path against the single-GPU result, as a callable.

Every rank computes the FULL problem on its own GPU (the single-process truth) and its SHARD through the sharded
entry points; the shard results must reproduce the matching rows / the global reductions:
interpolation parameter gradients (one packed all-reduce), DEC target distribution and fused KL step (global column
sum, global 'batchmean'), k-means (labels bit-exact, centres, inertia), the task-parallel and the row-sharded gap
sweep.  Used by tests/dist_gpu_check.py (torchrun, any world size) and by bench.py --gpus N, which prints the
report as the `parity` block of its JSON line so that a scaling record carries result checks, not only speed.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import functional as F_, parallel, synth
from .kmeans import KMeansB200


def run(rank, world, dev):
    """Runs every check on an initialised NCCL process group; returns {name: achieved error}.  Raises AssertionError
    on a mismatch."""
    import deep_interpolation_clustering_b200 as dic
    report = {}

    # ---- interpolation: gradients of the sharded batch == gradients of the whole batch ----------------------
    B, C, T, R, H = 4096, 6, 64, 48, 24.0
    xn = synth.make_encounters(B, C, T, H, seed=0)
    p = synth.make_interp_params(C, seed=1)
    rng = np.random.RandomState(2)
    vn = rng.normal(size=(B, C, R)).astype(np.float32)
    gn = rng.normal(size=(B, R, 3 * C)).astype(np.float32)

    def grads_of(sl):
        sci = dic.SingleChannelInterp(R, H, C, T, dev)
        cci = dic.CrossChannelInterp(C, T, dev)
        rbf = dic.RBF(H, R, C, C, 0.0, dic.basis_func_dict()["gaussian"], dev)
        rbf.compress_fc = torch.nn.Identity()
        sci.kernel.data = torch.tensor(p["sci_kernel"], device=dev)
        cci.kernel.data = torch.tensor(p["cci_kernel"], device=dev)
        rbf.kernel.data = torch.tensor(p["rbf_kernel"], device=dev)
        x = torch.tensor(xn[sl], device=dev)
        v = torch.tensor(vn[sl], device=dev)
        out = cci(sci(x))
        rec = rbf(v, x)
        ((out * torch.tensor(gn[sl], device=dev)).sum() + (rec ** 2).sum()).backward()
        return [sci.kernel, cci.kernel, rbf.kernel]

    full = grads_of(slice(0, B))
    lo, hi = parallel.shard_range(B, rank, world)
    mine = grads_of(slice(lo, hi))
    parallel.allreduce_gradients(mine)
    for name, a, b in zip(("d_sci", "d_cci", "d_rbf"), full, mine):
        err = float(((a.grad - b.grad).abs() / (a.grad.abs() + 1e-3 * a.grad.abs().max())).max())
        report[name] = err
        assert err < 2e-5, (name, err)          # different float32 summation order only

    # ---- DEC: global target distribution and fused KL step ---------------------------------------------------
    N, D, K = 20000, 64, 4
    zn, mun = synth.make_latents(N, D, K, seed=3)
    z, mu = torch.tensor(zn, device=dev), torch.tensor(mun, device=dev)
    ref = F_.dec_kl_step(z, mu, 1.0, weight=10.0)
    lo, hi = parallel.shard_range(N, rank, world)
    sh = parallel.sharded_dec_kl_step(z[lo:hi].contiguous(), mu, 1.0, weight=10.0)
    assert torch.equal(sh["labels"], ref["labels"][lo:hi])
    report["dec_p"] = float((sh["p"] - ref["p"][lo:hi]).abs().max())
    report["dec_kl"] = float((sh["kl"] - ref["kl"]).abs().max() / ref["kl"].abs().max())
    report["dec_dmu"] = float((sh["grad_mu"] - ref["grad_mu"]).abs().max() / ref["grad_mu"].abs().max())
    report["dec_dz"] = float((sh["grad_z"] - ref["grad_z"][lo:hi]).abs().max() / ref["grad_z"].abs().max())
    assert report["dec_p"] < 1e-6 and report["dec_kl"] < 1e-6 and report["dec_dmu"] < 1e-5 and report["dec_dz"] < 1e-5, report
    q = F_.dec_soft_assign(z, mu, 1.0).detach()
    p_full = F_.dec_target_distribution(q)
    p_sh = parallel.sharded_target_distribution(q[lo:hi].contiguous())
    report["target_p"] = float((p_sh - p_full[lo:hi]).abs().max())
    assert report["target_p"] < 1e-6

    # ---- k-means: sharded fit == single-process fit ----------------------------------------------------------
    for dtype in (np.float32, np.float64):
        X = synth.make_blobs(30001, 64, 5, seed=4).astype(dtype)          # odd size: ragged shards
        single = KMeansB200(n_clusters=5, n_init=2, random_state=7, device=dev).fit(X)
        lo, hi = parallel.shard_range(X.shape[0], rank, world)
        shard = KMeansB200(n_clusters=5, n_init=2, random_state=7, device=dev, sharded=True).fit(X[lo:hi])
        assert np.array_equal(shard.labels_, single.labels_[lo:hi]), "sharded k-means labels differ"
        tag = "f32" if dtype == np.float32 else "f64"
        report[f"km_{tag}_centers"] = float(np.abs(shard.cluster_centers_ - single.cluster_centers_).max())
        report[f"km_{tag}_inertia"] = abs(shard.inertia_ - single.inertia_) / single.inertia_
        assert report[f"km_{tag}_centers"] < (1e-5 if dtype == np.float32 else 1e-10), report
        assert report[f"km_{tag}_inertia"] < (1e-6 if dtype == np.float32 else 1e-12), report
        assert shard.n_iter_ == single.n_iter_

    # ---- gap sweep: (k, reference set) tasks dealt to the ranks == the same tasks on one GPU -----------------
    from .gap import KM
    Xg = synth.make_blobs(20000, 64, 4, seed=6).astype(np.float32)
    args = dict(k_max=6, n_references=3, version=1, draw="device", seed=5)
    solo = [dist.new_group([r]) for r in range(world)][rank]         # a world of one: every task on this GPU
    one = KM(6, internal_metrics=["Calinski-Harabasz"]).compute_gap_internal_metric(
        KMeansB200(n_init=3, random_state=9, device=dev), Xg, group=solo, **args)
    many = KM(6, internal_metrics=["Calinski-Harabasz"]).compute_gap_internal_metric(
        KMeansB200(n_init=3, random_state=9, device=dev), Xg, group=dist.group.WORLD, **args)
    a, b = one.astype(float).to_numpy(), many.astype(float).to_numpy()
    report["gap_tasks_max_diff"] = float(np.abs(a - b).max())
    assert np.allclose(a, b, rtol=1e-12, atol=0), (a, b)
    assert int(many["k"][many["gap"].astype(float).idxmax()]) == 4

    # ---- gap sweep, rows sharded: sharded fits + striped pairwise inertia == the whole matrix on one GPU ----
    class Sliced:                                  # t-th call -> rows [lo, hi) of the t-th pre-drawn uniform matrix
        def __init__(self, U, lo, hi):
            self.U, self.lo, self.hi, self.t = U, lo, hi, 0

        def __call__(self, shape):
            self.t += 1
            return self.U[self.t - 1][self.lo:self.hi]

    Xr = synth.make_blobs(6001, 64, 3, seed=8).astype(np.float32)
    U = np.random.RandomState(4).random_sample((3 * 2,) + Xr.shape)
    lo, hi = parallel.shard_range(Xr.shape[0], rank, world)
    whole = KM(4).compute_gap_internal_metric(KMeansB200(n_init=2, random_state=5, device=dev), Xr, k_max=4,
                                              n_references=2, version=1, draw=Sliced(U, 0, Xr.shape[0]),
                                              task_parallel=False)
    rows = KM(4).compute_gap_internal_metric(KMeansB200(n_init=2, random_state=5, device=dev, sharded=True),
                                             Xr[lo:hi], k_max=4, n_references=2, version=1, draw=Sliced(U, lo, hi),
                                             row_sharded=True, group=dist.group.WORLD)
    a, b = whole.astype(float).to_numpy(), rows.astype(float).to_numpy()
    report["gap_rows_max_rel_diff"] = float((np.abs(a - b) / np.maximum(np.abs(a), 1e-12)).max())
    assert np.allclose(a, b, rtol=1e-5, atol=1e-6), (a, b)

    return report
