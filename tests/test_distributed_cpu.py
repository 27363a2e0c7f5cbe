"""world_size-2 gloo tests (CPU) of the multi-GPU host logic.

The exchange steps live in deep_interpolation_clustering_b200/parallel.py and in
KMeansB200(sharded=True).  The CUDA kernels are replaced by oracle-backed stand-ins (injected
through the ``*_fn`` / ``_backend`` hooks) so that what is tested here is exactly the
sharding arithmetic and the collectives: every rank, fed its row shard, must reproduce the
single-process result on the concatenated batch.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

WORLD = 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(fn, *args):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_entry, args=(fn, r, port, q) + args) for r in range(WORLD)]
    for p in procs:
        p.start()
    results = {}
    for _ in range(WORLD):                       # drain before join: a full pipe would block the child
        r, res = q.get()
        results[r] = res
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0, f"rank exited with {p.exitcode}"
    for r, res in results.items():
        assert not isinstance(res, str) or not res.startswith("ERROR"), res
    assert sorted(results) == list(range(WORLD))
    return results


def _plain(obj):
    """Tensors do not survive the child's exit when sent through a queue: ship numpy instead."""
    if isinstance(obj, torch.Tensor):
        return obj.detach().cpu().numpy()
    if isinstance(obj, dict):
        return {k: _plain(v) for k, v in obj.items()}
    if isinstance(obj, (tuple, list)):
        return type(obj)(_plain(v) for v in obj)
    return obj


def _entry(fn, rank, port, q, *args):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(WORLD))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        q.put((rank, _plain(fn(rank, *args))))
    except Exception:                                # report instead of hanging the parent
        import traceback
        q.put((rank, "ERROR " + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


# ---- oracle-backed stand-ins for the device kernels ------------------------------------------
def _colsum(q):
    return q.double().sum(0)


def _p_from(q, f):
    from oracle import dec_oracle
    return torch.from_numpy(dec_oracle.target_distribution(q.numpy().astype(np.float64), colsum=f.numpy()))


def _assign(z, mu, alpha):
    from oracle import dec_oracle
    q = torch.from_numpy(dec_oracle.soft_assign(z.numpy().astype(np.float64), mu.numpy().astype(np.float64), alpha))
    return dict(q=q, labels=q.argmax(1).int(), colsum=q.sum(0))


def _kl_from(z, mu, f, alpha, weight=1.0, batch=None):
    from oracle import dec_oracle
    zn, mun = z.numpy().astype(np.float64), mu.numpy().astype(np.float64)
    q = dec_oracle.soft_assign(zn, mun, alpha)
    p = dec_oracle.target_distribution(q, colsum=f.numpy())
    dz, dmu = dec_oracle.kl_backward_closed_form(zn, mun, p, alpha, batch=batch, weight=weight)
    kl = weight * dec_oracle.kl_div_batchmean(p, q, batch=batch)
    return dict(p=torch.from_numpy(p), kl=torch.tensor([kl], dtype=torch.float64), grad_z=torch.from_numpy(dz),
                grad_mu=torch.from_numpy(dmu))


class _CpuBackend:
    """numpy stand-in for kmeans._Device with the same assign / min_d2 contract."""

    def __init__(self, X, K):
        self.X, self.N, self.D, self.K, self.dev = X, X.shape[0], X.shape[1], K, X.device
        self.labels = torch.full((self.N,), -1, dtype=torch.int32)
        self.sums = torch.zeros((K, self.D), dtype=torch.float64)
        self.counts = torch.zeros(K, dtype=torch.float64)
        self.stats = torch.zeros(4, dtype=torch.float64)

    def assign(self, centers, flags=0, want_sums=True, labels=None):
        from oracle import kmeans_oracle
        X, C = self.X.numpy(), centers.numpy()
        old = self.labels.numpy().copy()
        lab = old if flags & 2 else kmeans_oracle.e_step(X, C)
        self.labels.copy_(torch.from_numpy(lab))
        d2 = ((X.astype(np.float64) - C.astype(np.float64)[lab]) ** 2).sum(1)
        self.stats[:] = torch.tensor([d2.sum(), float((lab != old).sum()) if flags & 1 else 0.0,
                                      np.sqrt(d2).sum(), 0.0])
        if want_sums:
            s = np.zeros((centers.shape[0], self.D))
            np.add.at(s, lab, X.astype(np.float64))
            self.sums.copy_(torch.from_numpy(s))
            self.counts.copy_(torch.from_numpy(np.bincount(lab, minlength=centers.shape[0]).astype(np.float64)))

    def min_d2(self, cands, prev, out):
        d = ((self.X[None].double() - cands[:, None].double()) ** 2).sum(2)        # (L, N)
        if prev is not None:
            d = torch.minimum(d, prev.double()[None])
        if out is not None:
            out.copy_(d[0].to(out.dtype))
        return d.sum(1)


# ---- tests ------------------------------------------------------------------------------------
def _t_shard_and_grads(rank):
    from deep_interpolation_clustering_b200 import parallel
    assert parallel.world() == (rank, WORLD)
    spans = [parallel.shard_range(10, r, 3) for r in range(3)]
    assert spans == [(0, 4), (4, 7), (7, 10)]
    a = torch.nn.Parameter(torch.zeros(6))
    b = torch.nn.Parameter(torch.zeros(6, 6))
    c = torch.nn.Parameter(torch.zeros(3))                  # no grad: must be skipped
    a.grad = torch.full((6,), float(rank + 1))
    b.grad = torch.arange(36.0).view(6, 6) * (rank + 1)
    parallel.allreduce_gradients([a, b, c])
    return a.grad.clone(), b.grad.clone()


def test_shard_range_and_packed_gradient_allreduce():
    res = _run(_t_shard_and_grads)
    for r in range(WORLD):
        assert np.array_equal(res[r][0], np.full(6, 3.0, np.float32))
        assert np.array_equal(res[r][1], np.arange(36.0, dtype=np.float32).reshape(6, 6) * 3)


def _t_target_distribution(rank, q_all):
    from deep_interpolation_clustering_b200 import parallel
    lo, hi = parallel.shard_range(q_all.shape[0], rank, WORLD)
    return parallel.sharded_target_distribution(q_all[lo:hi], colsum_fn=_colsum, p_fn=_p_from)


def test_sharded_target_distribution_equals_global(golden):
    from oracle import dec_oracle
    g = golden("dec_k4")
    q = torch.from_numpy(g["q_f64"])
    res = _run(_t_target_distribution, q)
    got = np.concatenate([res[r] for r in range(WORLD)])
    assert np.allclose(got, g["p_f64"], rtol=1e-12, atol=0)
    # and it is NOT what per-shard normalisation (DataParallel's behaviour) would give
    per_shard = np.concatenate([dec_oracle.target_distribution(h) for h in (g["q_f64"][:32], g["q_f64"][32:])])
    assert not np.allclose(per_shard, g["p_f64"], rtol=1e-6)


def _t_dec_step(rank, z, mu):
    from deep_interpolation_clustering_b200 import parallel
    lo, hi = parallel.shard_range(z.shape[0], rank, WORLD)
    out = parallel.sharded_dec_kl_step(z[lo:hi], mu, 1.0, weight=10.0, assign_fn=_assign, kl_fn=_kl_from)
    return {k: out[k] for k in ("p", "kl", "grad_z", "grad_mu", "colsum", "batch")}


def test_sharded_dec_step_equals_global(golden):
    g = golden("dec_k4")
    res = _run(_t_dec_step, torch.from_numpy(g["z"]), torch.from_numpy(g["mu"]))
    assert res[0]["batch"] == res[1]["batch"] == g["z"].shape[0]
    p = np.concatenate([res[r]["p"] for r in range(WORLD)])
    gz = np.concatenate([res[r]["grad_z"] for r in range(WORLD)])
    assert np.allclose(p, g["p_f64"], rtol=1e-9, atol=1e-12)
    assert np.allclose(gz, 10.0 * g["dz_kl_f64"], rtol=1e-7, atol=1e-12)
    for r in range(WORLD):
        assert np.allclose(res[r]["kl"], 10.0 * g["kl_f64"], rtol=1e-9)
        assert np.allclose(res[r]["grad_mu"], 10.0 * g["dmu_kl_f64"], rtol=1e-6, atol=1e-10)
        assert np.allclose(res[r]["colsum"], g["q_f64"].sum(0), rtol=1e-9)


def _t_kmeans(rank, X, k, init):
    from deep_interpolation_clustering_b200 import parallel
    from deep_interpolation_clustering_b200.kmeans import KMeansB200
    lo, hi = parallel.shard_range(X.shape[0], rank, WORLD)
    km = KMeansB200(n_clusters=k, init=init, n_init=1 if init is not None and not isinstance(init, str) else 2,
                    random_state=11, sharded=True, _backend=_CpuBackend).fit(X[lo:hi])
    return km.labels_, km.cluster_centers_, km.inertia_, km.n_iter_


@pytest.mark.parametrize("k", [4, 7])
def test_sharded_lloyd_equals_single_process(golden, k):
    """Row-sharded Lloyd with one packed all-reduce per iteration == sklearn on the whole matrix."""
    g = golden("kmeans")
    X = g["X"].astype(np.float64)
    res = _run(_t_kmeans, X, k, X[:k].copy())
    labels = np.concatenate([res[r][0] for r in range(WORLD)])
    assert np.array_equal(labels, g[f"f64_k{k}_labels"])
    for r in range(WORLD):
        assert np.allclose(res[r][1], g[f"f64_k{k}_centers"], rtol=1e-9, atol=1e-9)
        assert np.isclose(res[r][2], float(g[f"f64_k{k}_inertia"]), rtol=1e-9)
        assert res[r][3] == int(g[f"f64_k{k}_n_iter"])


def test_sharded_relocation_and_kmeanspp(golden):
    g = golden("kmeans")
    X = g["X"].astype(np.float64)
    res = _run(_t_kmeans, X, 3, g["reloc_init"].astype(np.float64))      # empty cluster in iteration 1
    labels = np.concatenate([res[r][0] for r in range(WORLD)])
    assert np.array_equal(labels, g["reloc_labels"])
    assert np.allclose(res[0][1], g["reloc_centers"], rtol=1e-5, atol=1e-5)
    # k-means++ with a shared seed: both ranks agree, and match the unsharded run of the same code
    from deep_interpolation_clustering_b200.kmeans import KMeansB200
    res = _run(_t_kmeans, X, 5, "k-means++")
    assert np.allclose(res[0][1], res[1][1]) and res[0][2] == res[1][2]
    single = KMeansB200(n_clusters=5, n_init=2, random_state=11, _backend=_CpuBackend).fit(X)
    assert np.isclose(single.inertia_, res[0][2], rtol=1e-9)
    assert np.array_equal(np.concatenate([res[r][0] for r in range(WORLD)]), single.labels_)


def _cdist_sum(Xc, exact=False):
    return torch.cdist(Xc.double(), Xc.double()).sum().reshape(1)


def _gap_sweep(group_mode):
    from deep_interpolation_clustering_b200.gap import KM
    from deep_interpolation_clustering_b200.kmeans import KMeansB200
    from deep_interpolation_clustering_b200 import synth
    X = synth.make_blobs(240, 8, 3, seed=5).astype(np.float32)
    km = KM(5, _pairwise=_cdist_sum, _device=torch.device("cpu"))
    clustering = KMeansB200(n_init=2, random_state=3, _backend=_CpuBackend)
    df = km.compute_gap_internal_metric(clustering, X, k_max=5, n_references=3, version=1, seed=7, **group_mode)
    return df.astype(float).to_numpy()


def _t_gap_tasks(rank):
    return _gap_sweep({"group": dist.group.WORLD})


def test_task_parallel_gap_sweep_equals_single_process():
    """The (k, reference set) fits dealt round-robin to two ranks + one all-reduce of the inertia table give the
    single-process table exactly (per-task seeded reference draws, SURVEY 8e 'task-parallel at c4')."""
    res = _run(_t_gap_tasks)
    single = _gap_sweep({"task_parallel": True})
    assert np.isfinite(single).all() and single.shape == (4, 5)
    for r in range(WORLD):
        assert np.array_equal(res[r], single)
    assert int(single[np.argmax(single[:, 1]), 0]) == 3          # three blobs: the gap picks K = 3


def _cdist_sum_part(Xc, part, n_parts, exact=False):
    return torch.cdist(Xc[part::n_parts].double(), Xc.double()).sum().reshape(1)


class _SlicedDraws:
    """draw(shape) stand-in: the t-th call returns rows [lo, hi) of the t-th pre-generated uniform matrix."""

    def __init__(self, U, lo, hi):
        self.U, self.lo, self.hi, self.t = U, lo, hi, 0

    def __call__(self, shape):
        out = self.U[self.t][self.lo:self.hi]
        self.t += 1
        assert out.shape == tuple(shape)
        return out


def _gap_rows(rank, world):
    from deep_interpolation_clustering_b200 import parallel, synth
    from deep_interpolation_clustering_b200.gap import KM
    from deep_interpolation_clustering_b200.kmeans import KMeansB200
    X = synth.make_blobs(301, 8, 3, seed=5).astype(np.float32)                # odd size: ragged shards
    U = np.random.RandomState(3).random_sample((3 * 2,) + X.shape)
    lo, hi = parallel.shard_range(X.shape[0], rank, world) if world > 1 else (0, X.shape[0])
    km = KM(4, _pairwise=_cdist_sum, _pairwise_part=_cdist_sum_part, _device=torch.device("cpu"))
    clustering = KMeansB200(n_init=2, random_state=3, sharded=world > 1, _backend=_CpuBackend)
    df = km.compute_gap_internal_metric(clustering, X[lo:hi], k_max=4, n_references=2, version=1,
                                        draw=_SlicedDraws(U, lo, hi), row_sharded=world > 1,
                                        group=dist.group.WORLD if world > 1 else None,
                                        task_parallel=False)
    return df.astype(float).to_numpy()


def _t_gap_rows(rank):
    return _gap_rows(rank, WORLD)


def test_row_sharded_gap_sweep_equals_single_process():
    """Rows of the data and of every reference set sharded over two ranks: sharded k-means fits + striped pairwise
    inertia (all-gather of a cluster's rows, stripes of the tile list, one all-reduce) give the single-process table."""
    res = _run(_t_gap_rows)
    single = _gap_rows(0, 1)
    assert np.isfinite(single).all()
    for r in range(WORLD):
        assert np.allclose(res[r], single, rtol=1e-9, atol=1e-12), (res[r], single)
