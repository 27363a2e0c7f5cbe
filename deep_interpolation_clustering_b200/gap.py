"""K-selection sweep (gap statistic / elbow) on device - mirror of the ``KM`` class of
p2_clustering_optK.py:226-410 without its plotting.

``KM.compute_gap_internal_metric(clustering, data, k_max, n_references, version)`` keeps the
reference's signature and returns the same pandas DataFrame (index k = 2..k_max, columns
``['k','gap','ref','act','ref_s', *internal metric names]``).  The "inertia" of the reference
(mean intra-cluster mean pairwise Euclidean distance, :334-342) is evaluated by the tiled
``dic_pairwise_dist_sum`` kernel without materialising the n_c x n_c matrix, so the sweep
runs at N = 1M where the reference needs terabytes.

Reference draws use ``np.random.random_sample`` on the host exactly like :370 so that a
seeded run sees the same draws; pass ``draw=`` to inject your own.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import _lib
from . import internal_eval
from . import parallel
from .kmeans import KMeansB200, _Comm, _DT


def _as_device(X, device=None):
    if isinstance(X, torch.Tensor):
        if not X.is_cuda and device is None and torch.cuda.is_available():
            X = X.cuda()
        return X.contiguous()
    X = np.asarray(X)
    if X.dtype not in (np.float32, np.float64):
        X = X.astype(np.float64)
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    return torch.from_numpy(np.ascontiguousarray(X)).to(dev)


def _data_range(data):
    """(min, max - min) of the data matrix as numpy scalars in the data's own dtype: p2_clustering_optK.py:360
    (`data.min()`, `data.max() - data.min()` on the float32 feature matrix)."""
    if isinstance(data, torch.Tensor):
        mn, mx = data.min().cpu().numpy()[()], data.max().cpu().numpy()[()]
    else:
        mn, mx = data.min(), data.max()
    return mn, mx - mn


TC_MIN_ROWS = 512      # below this the exact CUDA-core kernel is used and the dtype is preserved
DIC_PAIRWISE_EXACT = 16  # include/dic_b200.h: OR-ed into dtype, forces the direct (x_i - x_j)^2 kernel


def _pairwise_call(fn_name, Xc, exact, out, ws, extra=()):
    """Shared body of pairwise_dist_sum / pairwise_dist_sum_part.  Asynchronous: the NaN the tensor-core kernels
    write when their pipeline stalls is checked by the caller when it reads the value (check_pairwise_sums)."""
    if not exact and Xc.shape[0] >= TC_MIN_ROWS and Xc.shape[1] % 4 == 0:
        # distances are translation invariant; the Gram-form tensor-core kernel wants ||x||^2 small against them
        Xc = (Xc - Xc.mean(dim=0, keepdim=True)).to(torch.float32)
    Xc = Xc.contiguous()
    if Xc.data_ptr() % 16:
        Xc = Xc.clone()
    n, D = Xc.shape
    if out is None:
        out = torch.empty(1, dtype=torch.float64, device=Xc.device)
    L = _lib.lib()
    need = int(L.dic_pairwise_workspace_bytes(n, D))
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=Xc.device)
    dt = _DT[Xc.dtype] | (DIC_PAIRWISE_EXACT if exact else 0)
    with torch.cuda.device(Xc.device):
        _lib.check(getattr(L, fn_name)(_lib.ptr(Xc), _lib.ptr(out), _lib.ptr(ws), n, D, dt, *extra,
                                       _lib.current_stream(Xc.device)), fn_name)
    return out


def check_pairwise_sums(values, what="dic_pairwise_dist_sum"):
    """values: host floats read back from the kernels.  NaN is the kernels' 'pipeline stalled' sentinel
    (pairwise_tc.cu: every mbarrier wait is bounded) - raise instead of letting it reach the gap table."""
    for v in values:
        if v != v:
            raise _lib.DicError(f"{what}: the tensor-core pipeline reported a stalled mbarrier hand-off (NaN result); "
                                "the launch did not complete its tile list")
    return values


def pairwise_dist_sum(Xc, exact=False, out=None, ws=None):
    """sum over the full n x n Euclidean distance matrix of the rows of Xc (device tensor) -> (1,) float64 tensor.

    Clusters of >= 512 rows are centred (distances are translation invariant) and evaluated in
    float32 on the tensor cores (tcgen05, split-operand products: float32-grade dot products, ~1e-7 relative on
    the sum); ``exact=True`` keeps the input dtype and ALWAYS takes the direct (x_i - x_j)^2 CUDA-core kernel
    (dtype | DIC_PAIRWISE_EXACT), whatever the cluster size.  ``out`` (1-element float64 view) and ``ws`` (byte
    workspace) let a sweep reuse its buffers.
    """
    return _pairwise_call("dic_pairwise_dist_sum", Xc, exact, out, ws)


def pairwise_dist_sum_part(Xc, part, n_parts, exact=False, out=None, ws=None):
    """Stripe `part` of `n_parts` of pairwise_dist_sum(Xc): Xc holds ALL rows of the cluster (identical on every
    rank), the stripes tile the kernel's tile list once, so the n_parts results add up to the full sum."""
    return _pairwise_call("dic_pairwise_dist_sum_part", Xc, exact, out, ws, (int(part), int(n_parts)))


class KM(object):
    """p2_clustering_optK.py:226-410 (constructor signature kept; plots are out of scope)."""

    def __init__(self, k_max, out_path=None, internal_metrics=(), n_init=10, gap_b=10, exact_pairwise=False,
                 _pairwise=None, _device=None, _pairwise_part=None):
        self.exact_pairwise = exact_pairwise
        self._pairwise = _pairwise          # test hooks (gloo tests on CPU): distance-sum stand-ins and their device
        self._pairwise_part = _pairwise_part
        self._dev = _device
        self._comm = None                   # set while a row-sharded sweep runs
        self._ws = None                     # pairwise workspace, reused across the evaluations of a sweep
        self.timers = None                  # set to {} to collect per-phase seconds of a task-parallel sweep (bench.py)
        self.k_max = k_max
        self.out_path = os.path.join(out_path, "plot") if out_path else None
        if self.out_path:
            os.makedirs(self.out_path, exist_ok=True)
        self.internal_metrics_names = list(internal_metrics)
        self.internal_metrics = self.get_internal_metrics()
        self.n_init = n_init
        self.gap_b = gap_b

    def get_internal_metrics(self):
        table = {"Dunn_Index": internal_eval.DunnIndex, "Sihouette": internal_eval.Sihouette,
                 "Davies-Bouldin_Index": internal_eval.DBIndex, "Calinski-Harabasz": internal_eval.CHIndex}
        return [table[name]() for name in self.internal_metrics_names]

    def _timed(self, key, fn, *args):
        """fn(*args); with self.timers set, bracketed by device synchronisations and accumulated under `key`."""
        if self.timers is None:
            return fn(*args)
        import time
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn(*args)
        torch.cuda.synchronize()
        self.timers[key] = self.timers.get(key, 0.0) + time.perf_counter() - t0
        self.timers[key + "_calls"] = self.timers.get(key + "_calls", 0) + 1
        return out

    # ---- the two "inertia" definitions ---------------------------------------------------------
    def _cluster_sums(self, a, X):
        """[(sum of the full n_c x n_c distance matrix, n_c)] for the labels present in `a` (np.unique order).
        The rows are sorted by label ONCE (one stable argsort + one gather), the per-cluster launches share one
        workspace and write into one (K,) buffer, and the host reads that buffer once: two synchronisations per
        evaluation instead of two per cluster."""
        if self._comm is not None and self._comm.on:
            return self._cluster_sums_sharded(a, X, self._comm)
        Xd = _as_device(X, self._dev)
        ad = torch.as_tensor(np.asarray(a) if not isinstance(a, torch.Tensor) else a).to(Xd.device).long()
        if ad.numel() == 0:
            return []
        order = torch.argsort(ad, stable=True)
        Xs = Xd[order]
        counts = torch.bincount(ad).tolist()                     # sync 1: cluster sizes
        pw = self._pairwise or pairwise_dist_sum
        sums = torch.zeros(len(counts), dtype=torch.float64, device=Xd.device)
        if self._pairwise is None:
            need = int(_lib.lib().dic_pairwise_workspace_bytes(max(counts), Xd.shape[1]))
            if self._ws is None or self._ws.numel() < need or self._ws.device != Xd.device:
                self._ws = torch.empty(need, dtype=torch.uint8, device=Xd.device)
        off = 0
        for c, n in enumerate(counts):
            if n:
                Xc = Xs[off:off + n]
                if self._pairwise is None:
                    pw(Xc, exact=self.exact_pairwise, out=sums[c:c + 1], ws=self._ws)
                else:
                    sums[c:c + 1] = pw(Xc, exact=self.exact_pairwise)
                off += n
        host = check_pairwise_sums(sums.tolist())                # sync 2: the K sums
        return [(host[c], n) for c, n in enumerate(counts) if n]

    def _cluster_sums_sharded(self, a, X, comm):
        """_cluster_sums for rows sharded over the ranks of `comm` (pairwise inertia): the rows of one
        cluster are all-gathered (padded to the largest shard), every rank evaluates its stripe of the tile list
        (dic_pairwise_dist_sum_part) and ONE all-reduce of the K per-cluster sums ends the evaluation."""
        Xd = _as_device(X, self._dev)
        ad = torch.as_tensor(np.asarray(a) if not isinstance(a, torch.Tensor) else a).to(Xd.device).long()
        top = torch.tensor([int(ad.max()) + 1 if ad.numel() else 0], dtype=torch.int64, device=Xd.device)
        K = int(comm.gather(top).max())
        counts = comm.gather(torch.bincount(ad, minlength=K)).cpu()          # (ranks, K)
        total = counts.sum(0)
        pw = self._pairwise_part or pairwise_dist_sum_part
        sums = torch.zeros(K, dtype=torch.float64, device=Xd.device)
        for c in range(K):
            if int(total[c]) == 0:
                continue
            mine = Xd[ad == c]
            pad = torch.zeros((int(counts[:, c].max()), Xd.shape[1]), dtype=Xd.dtype, device=Xd.device)
            pad[:mine.shape[0]] = mine
            parts = comm.gather(pad)
            Xc = torch.cat([parts[r, :int(counts[r, c])] for r in range(comm.size)])
            sums[c:c + 1] = pw(Xc, comm.rank, comm.size, exact=self.exact_pairwise)
        if bool(torch.isnan(sums).any()):          # a stalled pipeline must not ride through the all-reduce as NaN
            check_pairwise_sums([float("nan")], "dic_pairwise_dist_sum_part")
        comm.sum_(sums)
        host = sums.tolist()
        return [(host[c], int(total[c])) for c in range(K) if int(total[c]) > 0]

    def compute_inertia_v1(self, a, X):
        """mean_c [ mean of the full n_c x n_c distance matrix ]   (:334-342)."""
        W = [float(s) / (n * n) for s, n in self._cluster_sums(a, X)]
        return float(np.mean(W))

    def computer_intertia_v2(self, a, X):
        """sum_c [ sum of the full distance matrix / (2 n_c) ]      (:344-351, name as upstream)."""
        return float(sum(float(s) / (2 * n) for s, n in self._cluster_sums(a, X)))

    # ---- gap statistic -----------------------------------------------------------------------
    def compute_gap_internal_metric(self, clustering, data, k_max=5, n_references=5, version=2, draw=None,
                                    group=None, task_parallel=None, seed=0, row_sharded=False):
        """``row_sharded=True``: ``data`` holds this rank's ROWS only and ``clustering`` is a
        ``KMeansB200(sharded=True, process_group=group)``; the data range is all-reduced, every rank draws its own rows
        of each reference set, the fits exchange one packed all-reduce per Lloyd iteration and the pairwise inertia
        runs in stripes (_cluster_sums_sharded) - the mode for matrices beyond one GPU (SURVEY 8e "row-shard at c5").
        ``group`` / ``task_parallel=True``: the (k, reference set) fits of the sweep are independent, so they are
        dealt round-robin to the ranks of ``group`` (every rank holds the whole data matrix, SURVEY 8e "task-parallel
        at c4") and ONE all-reduce of the (k, reference) table of inertias ends the sweep.  Reference set (k, j) is
        then drawn from its own generator seeded by (seed, k, j), so the table does not depend on the world size
        (give ``clustering`` an integer ``random_state`` for the same property of its k-means++ draws)."""
        import pandas as pd
        if not (isinstance(data, torch.Tensor) and data.is_cuda and _accepts_tensor(clustering)):
            # host arrays as the reference passes them; a CUDA tensor stays where it is for a device estimator
            data = np.asarray(data) if not isinstance(data, torch.Tensor) else data.cpu().numpy()
        if len(data.shape) == 1:
            data = data.reshape(-1, 1)
        if row_sharded:
            return self._gap_row_sharded(clustering, data, k_max, n_references, version, draw, group, seed)
        if task_parallel or (task_parallel is None and group is not None):
            return self._gap_task_parallel(clustering, data, k_max, n_references, version, draw, group, seed)
        # draw: None = np.random.random_sample like the reference (:370; host RNG, exact stream parity);
        # "device" = uniform float64 draws generated on the GPU (same distribution, no 8*N*D-byte host
        # generation + upload per reference set - at 1M x 64 that is 0.5 s of host time 180 times);
        # "device32" = the same in float32 (the reference's float64 draws make sklearn - and the Lloyd kernels here -
        # iterate in float64 on the reference sets although the data is float32; opt-in, half the bytes per pass)
        device_draws = isinstance(draw, str) and draw in ("device", "device32")
        ref_dtype = torch.float32 if draw == "device32" else torch.float64
        draw = np.random.random_sample if (draw is None or device_draws) else draw
        inertia = self.compute_inertia_v1 if version == 1 else self.computer_intertia_v2
        data_min, data_rng = _data_range(data)                                       # :360
        k_rng = range(2, k_max + 1)
        vals = pd.DataFrame(index=k_rng, columns=["k", "gap", "ref", "act", "ref_s"] + self.internal_metrics_names)
        data_dev = _as_device(data, self._dev)
        for k in k_rng:
            local_inertia = []
            clustering.n_clusters = k                                                # :367
            for _ in range(n_references):
                if device_draws and _accepts_tensor(clustering):
                    ref_dev = torch.rand(data.shape, dtype=ref_dtype, device=data_dev.device) * float(data_rng) \
                        + float(data_min)
                    reference = None
                else:
                    reference = draw(data.shape) * data_rng + data_min               # :370 (float64)
                    ref_dev = _as_device(reference, self._dev)
                assignments = clustering.fit_predict(ref_dev if _accepts_tensor(clustering) else reference)
                local_inertia.append(inertia(assignments, ref_dev))
            ref = np.mean(np.log(local_inertia))                                     # :374
            ref_s = np.sqrt(1 + 1 / n_references) * np.std(np.log(local_inertia))    # :375
            assignments = clustering.fit_predict(data_dev if _accepts_tensor(clustering) else data)
            act = np.log(inertia(assignments, data_dev))                             # :377-379
            gap = ref - act                                                          # :393
            a_host = assignments.cpu().numpy() if isinstance(assignments, torch.Tensor) else np.asarray(assignments)
            metric_values = [m(data_dev, a_host) for m in self.internal_metrics]     # :401-405
            vals.loc[k] = [k, gap, ref, act, ref_s] + metric_values
        return vals

    def _gap_task_parallel(self, clustering, data, k_max, n_references, version, draw, group, seed):
        """The sweep above as a flat list of (k, j) tasks, j < n_references a reference set and j = n_references the
        data itself (with its internal metrics), dealt round-robin to the ranks; see compute_gap_internal_metric."""
        import pandas as pd
        import torch.distributed as dist
        rank, ws = parallel.world(group)
        device_draws = isinstance(draw, str) and draw in ("device", "device32")
        ref_dtype = torch.float32 if draw == "device32" else torch.float64
        inertia = self.compute_inertia_v1 if version == 1 else self.computer_intertia_v2
        data_min, data_rng = _data_range(data)
        k_rng = range(2, k_max + 1)
        n_m = len(self.internal_metrics)
        table = np.zeros((len(k_rng), n_references + 1 + n_m), dtype=np.float64)
        data_dev = _as_device(data, self._dev)
        on_dev = _accepts_tensor(clustering)
        task = 0
        for ki, k in enumerate(k_rng):
            for j in range(n_references + 1):
                mine = task % ws == rank
                task += 1
                if not mine:
                    continue
                clustering.n_clusters = k
                if j == n_references:
                    a = self._timed("fit_data", clustering.fit_predict, data_dev if on_dev else data)
                    table[ki, j] = self._timed("inertia_data", inertia, a, data_dev)
                    a_host = a.cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
                    table[ki, j + 1:] = [self._timed("metrics", m, data_dev, a_host) for m in self.internal_metrics]
                    continue
                task_seed = (int(seed) * 1000003 + k * 1009 + j) % (2 ** 31)
                if device_draws and on_dev:
                    def draw_dev():
                        gen = torch.Generator(device=data_dev.device).manual_seed(task_seed)
                        return torch.rand(data.shape, dtype=ref_dtype, device=data_dev.device, generator=gen) \
                            * float(data_rng) + float(data_min)
                    ref_dev = self._timed("draw", draw_dev)
                    reference = None
                else:
                    sample = np.random.RandomState(task_seed).random_sample if (draw is None or device_draws) else draw
                    reference = sample(data.shape) * data_rng + data_min
                    ref_dev = _as_device(reference, self._dev)
                a = self._timed("fit_ref", clustering.fit_predict, ref_dev if on_dev else reference)
                table[ki, j] = self._timed("inertia_ref", inertia, a, ref_dev)
        if ws > 1:
            nccl = dist.get_backend(group) == "nccl"
            buf = torch.from_numpy(table).to(data_dev.device) if nccl else torch.from_numpy(table)
            dist.all_reduce(buf, group=group)                    # every cell was written by exactly one rank
            table = buf.cpu().numpy()
        vals = pd.DataFrame(index=k_rng, columns=["k", "gap", "ref", "act", "ref_s"] + self.internal_metrics_names)
        for ki, k in enumerate(k_rng):
            logs = np.log(table[ki, :n_references])
            ref = np.mean(logs)
            ref_s = np.sqrt(1 + 1 / n_references) * np.std(logs)
            act = np.log(table[ki, n_references])
            vals.loc[k] = [k, ref - act, ref, act, ref_s] + list(table[ki, n_references + 1:])
        return vals

    def _gap_row_sharded(self, clustering, data, k_max, n_references, version, draw, group, seed):
        import pandas as pd
        import torch.distributed as dist
        comm = _Comm(group)
        device_draws = isinstance(draw, str) and draw in ("device", "device32")
        ref_dtype = torch.float32 if draw == "device32" else torch.float64
        inertia = self.compute_inertia_v1 if version == 1 else self.computer_intertia_v2
        data_dev = _as_device(data, self._dev)
        lohi = torch.tensor([float(data.min()), -float(data.max())], dtype=torch.float64, device=data_dev.device)
        if comm.on:
            dist.all_reduce(lohi, op=dist.ReduceOp.MIN, group=group)
        np_t = np.float32 if (data.dtype in (np.float32, torch.float32)) else np.float64
        data_min = np_t(float(lohi[0]))                                             # :360 over ALL rows, in the
        data_rng = np_t(float(-lohi[1])) - data_min                                 # data's own arithmetic
        k_rng = range(2, k_max + 1)
        vals = pd.DataFrame(index=k_rng, columns=["k", "gap", "ref", "act", "ref_s"] + self.internal_metrics_names)
        on_dev = _accepts_tensor(clustering)
        self._comm = comm
        try:
            for k in k_rng:
                local_inertia = []
                clustering.n_clusters = k
                for j in range(n_references):
                    task_seed = ((int(seed) * 1000003 + k * 1009 + j) * 1031 + comm.rank) % (2 ** 31)
                    if device_draws and on_dev:
                        gen = torch.Generator(device=data_dev.device).manual_seed(task_seed)
                        ref_dev = torch.rand(data.shape, dtype=ref_dtype, device=data_dev.device, generator=gen) \
                            * float(data_rng) + float(data_min)
                        reference = None
                    else:
                        sample = np.random.RandomState(task_seed).random_sample if (draw is None or device_draws) \
                            else draw
                        reference = sample(data.shape) * data_rng + data_min
                        ref_dev = _as_device(reference, self._dev)
                    a = clustering.fit_predict(ref_dev if on_dev else reference)
                    local_inertia.append(inertia(a, ref_dev))
                ref = np.mean(np.log(local_inertia))
                ref_s = np.sqrt(1 + 1 / n_references) * np.std(np.log(local_inertia))
                a = clustering.fit_predict(data_dev if on_dev else data)
                act = np.log(inertia(a, data_dev))
                # :401-405 on the sharded rows: CH / DB all-reduce K-sized statistics, the O(N^2) scores gather the rows
                metric_values = [m(data_dev, a, group=group) for m in self.internal_metrics]
                vals.loc[k] = [k, ref - act, ref, act, ref_s] + metric_values
        finally:
            self._comm = None
        return vals

    # ---- elbow ---------------------------------------------------------------------------------
    def elbow(self, train_feat, valid_feat, device=None):
        """p2_clustering_optK.py:255-265: distortion = sum_i min_j ||x_i - c_j|| / N per k."""
        tr, va = [], []
        Xt, Xv = _as_device(train_feat, device), _as_device(valid_feat, device)
        for k in range(2, self.k_max + 1):
            km = KMeansB200(n_clusters=k, init="k-means++").fit(Xt)                  # :260
            tr.append(km.score_distortion(Xt))
            va.append(km.score_distortion(Xv))
        return tr, va

    def train(self, train_data, valid_data, select_opt_k, **kwargs):
        """Numbers of p2_clustering_optK.py:250-332 (CSV written when out_path is set; no plots)."""
        train_feat, valid_feat = train_data["hidden"], valid_data["hidden"]
        results = {}
        for method in select_opt_k:
            if method == "elbow":
                tr, va = self.elbow(train_feat, valid_feat)
                results["elbow"] = (tr, va)
                if self.out_path:      # the numbers behind {train,valid}_elbow.png (:266-274), one row per k
                    import pandas as pd
                    pd.DataFrame({"k": list(range(2, self.k_max + 1)), "train_distortion": tr,
                                  "valid_distortion": va}).to_csv(os.path.join(self.out_path, "elbow.csv"), index=False)
            elif method == "gap_sts":
                df = self.compute_gap_internal_metric(KMeansB200(n_init=self.n_init), train_feat, self.k_max,
                                                      n_references=self.gap_b, version=1).astype(float)
                if self.out_path:
                    df.to_csv(os.path.join(self.out_path, "gap_sts_v1.csv"), index=False)
                results["gap_sts"] = df
        return results


def _accepts_tensor(clustering):
    return isinstance(clustering, KMeansB200) or getattr(clustering, "accepts_device_tensors", False)
