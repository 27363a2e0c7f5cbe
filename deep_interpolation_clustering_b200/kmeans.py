"""sklearn-``KMeans``-compatible estimator backed by the sm_100a Lloyd kernels.

Stands behind the reference's calls ``KMeans(n_clusters=K, n_init=20).fit_predict(hidden)``
(clustering_trainer.py:75-82), ``KMeans(n_init=...)`` handed to the gap routine
(p2_clustering_optK.py:284, with ``clustering.n_clusters = k`` set at :367),
``KMeans(n_clusters=k, init='k-means++').fit(train_feat)`` (:260) and
p4_clustering_final.py:159-174.  Same duck type: constructor kwargs ``n_clusters, init,
n_init, max_iter, tol, random_state``; settable ``n_clusters``; ``fit / fit_predict /
predict``; ``cluster_centers_`` (writable ndarray), ``labels_``, ``inertia_``, ``n_iter_``.
Inputs and outputs are host numpy arrays (CUDA tensors are accepted and stay on device).

The algorithm follows scikit-learn 1.9.0 step by step: dtype preserved
(float32 data, float64 reference draws), mean-centring, tol * mean(var), k-means++ with
2+int(ln K) greedy trials consuming the numpy RandomState exactly like sklearn (so a shared
global stream stays in step), strict-'<' argmin, empty-cluster relocation, label-equality or
centre-shift stopping, a final E-step when not strictly converged, best of n_init by inertia.
"""
from __future__ import annotations

import numbers

import numpy as np
import torch
import torch.distributed as dist

from . import _lib

_DT = {torch.float32: 0, torch.float64: 1}
DIC_KM_COUNT_CHANGES = 1
DIC_KM_KEEP_LABELS = 2
DIC_KM_NO_INERTIA = 4       # a Lloyd iteration does not need the inertia / distance sums


def _check_random_state(seed):
    """sklearn.utils.check_random_state: None -> numpy's global RandomState singleton."""
    if seed is None or seed is np.random:
        return np.random.mtrand._rand
    if isinstance(seed, numbers.Integral):
        return np.random.RandomState(seed)
    if isinstance(seed, np.random.RandomState):
        return seed
    raise ValueError(f"{seed!r} cannot be used to seed a numpy.random.RandomState instance")


def _same_clustering(a, b, K):
    """Equal up to a permutation of the labels (sklearn _k_means_common.pyx:_is_same_clustering)."""
    a = np.asarray(a, np.int64)
    b = np.asarray(b, np.int64)
    pairs = np.unique(a * K + b)
    return len(np.unique(pairs // K)) == len(pairs)


class _Comm:
    """The exchange steps of a row-sharded fit (identity when there is a single rank)."""

    def __init__(self, group):
        self.group = group
        self.on = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.rank = dist.get_rank(group) if self.on else 0
        self.size = dist.get_world_size(group) if self.on else 1

    def sum_(self, *tensors):
        """In-place all-reduce(sum) of several tensors as ONE packed float64 message."""
        if not self.on:
            return
        flat = torch.cat([t.reshape(-1).to(torch.float64) for t in tensors])
        dist.all_reduce(flat, group=self.group)
        off = 0
        for t in tensors:
            n = t.numel()
            t.copy_(flat[off:off + n].view_as(t).to(t.dtype))
            off += n

    def gather(self, t):
        """All-gather of equally shaped tensors -> stacked (size, ...)."""
        if not self.on:
            return t[None]
        out = [torch.empty_like(t) for _ in range(self.size)]
        dist.all_gather(out, t.contiguous(), group=self.group)
        return torch.stack(out)


class _Device:
    """Device-side state and kernel calls for one (X, K) problem."""

    def __init__(self, X, K):
        self.X = X
        self.N, self.D = X.shape
        self.K = K
        self.dt = _DT[X.dtype]
        self.dev = X.device
        L = _lib.lib()
        self.ws = torch.empty(max(int(L.dic_kmeans_workspace_bytes(max(K, 16), self.D)), 16), dtype=torch.uint8,
                              device=self.dev)
        self.labels = torch.full((self.N,), -1, dtype=torch.int32, device=self.dev)
        self.sums = torch.empty((K, self.D), dtype=torch.float64, device=self.dev)
        self.counts = torch.empty(K, dtype=torch.float64, device=self.dev)
        self.stats = torch.empty(4, dtype=torch.float64, device=self.dev)
        self.status = torch.empty(4, dtype=torch.float64, device=self.dev)
        self._step_args = None
        self._run_args = None
        self.status8 = None

    def lloyd_step(self, centers, flags):
        """One Lloyd iteration (E-step + M-step in place on `centers`) in ONE C call with ONE host sync;
        returns (changed, shift2, n_empty, inertia).  If a cluster came out empty the centres are untouched."""
        key = centers.data_ptr()
        if self._step_args is None or self._step_args[0] != key:
            L = _lib.lib()
            self._step_args = (key, L.dic_kmeans_lloyd_step,
                               (_lib.ptr(self.X), key, _lib.ptr(self.labels), _lib.ptr(self.sums), _lib.ptr(self.counts),
                                _lib.ptr(self.stats), _lib.ptr(self.status), _lib.ptr(self.ws), self.N, self.D,
                                centers.shape[0], self.dt), _lib.current_stream(self.dev))
        _, fn, args, stream = self._step_args
        if torch.cuda.current_device() == self.dev.index:
            rc = fn(*args, flags, stream)
        else:
            with torch.cuda.device(self.dev):
                rc = fn(*args, flags, stream)
        _lib.check(rc, "dic_kmeans_lloyd_step")
        return self.status.tolist()

    def lloyd_run(self, centers, flags, n_steps, tol):
        """Up to n_steps Lloyd iterations enqueued back to back with the stopping rule on the device
        (dic_kmeans_lloyd_run); ONE host sync.  Returns status8 = [changed, shift2, n_empty, inertia, stop flag,
        iterations so far, strict, -]."""
        key = centers.data_ptr()
        if self._run_args is None or self._run_args[0] != key:
            if self.status8 is None:
                self.status8 = torch.zeros(8, dtype=torch.float64, device=self.dev)
            self._run_args = (key, _lib.lib().dic_kmeans_lloyd_run,
                              (_lib.ptr(self.X), key, _lib.ptr(self.labels), _lib.ptr(self.sums), _lib.ptr(self.counts),
                               _lib.ptr(self.stats), _lib.ptr(self.status8), _lib.ptr(self.ws), self.N, self.D,
                               centers.shape[0], self.dt), _lib.current_stream(self.dev))
        _, fn, args, stream = self._run_args
        if torch.cuda.current_device() == self.dev.index:
            rc = fn(*args, flags, int(n_steps), float(tol), stream)
        else:
            with torch.cuda.device(self.dev):
                rc = fn(*args, flags, int(n_steps), float(tol), stream)
        _lib.check(rc, "dic_kmeans_lloyd_run")
        return self.status8.tolist()

    def update(self, centers):
        """M-step in place on `centers`; returns (changed, shift2, n_empty, inertia) with ONE sync."""
        with torch.cuda.device(self.dev):
            _lib.check(_lib.lib().dic_kmeans_update(
                _lib.ptr(self.sums), _lib.ptr(self.counts), _lib.ptr(self.stats), _lib.ptr(centers),
                _lib.ptr(self.status), self.D, centers.shape[0], self.dt, _lib.current_stream(self.dev)),
                "dic_kmeans_update")
        return self.status.tolist()

    def assign(self, centers, flags=0, want_sums=True, labels=None):
        labels = self.labels if labels is None else labels
        with torch.cuda.device(self.dev):
            _lib.check(_lib.lib().dic_kmeans_assign(
                _lib.ptr(self.X), _lib.ptr(centers), _lib.ptr(labels),
                _lib.ptr(self.sums) if want_sums else None, _lib.ptr(self.counts) if want_sums else None,
                _lib.ptr(self.stats), _lib.ptr(self.ws), self.N, self.D, centers.shape[0], self.dt, flags,
                _lib.current_stream(self.dev)), "dic_kmeans_assign")

    def min_d2(self, cands, prev, out):
        L = cands.shape[0]
        pots = torch.empty(L, dtype=torch.float64, device=self.dev)
        with torch.cuda.device(self.dev):
            _lib.check(_lib.lib().dic_kmeans_min_d2(
                _lib.ptr(self.X), _lib.ptr(cands), _lib.ptr(prev), _lib.ptr(out), _lib.ptr(pots),
                _lib.ptr(self.ws), self.N, self.D, L, self.dt, _lib.current_stream(self.dev)),
                "dic_kmeans_min_d2")
        return pots


class KMeansB200:
    """K-means on a B200; drop-in for ``sklearn.cluster.KMeans`` as the reference uses it.

    ``process_group``: when given (or when torch.distributed is initialised and
    ``sharded=True``), X holds only this rank's ROWS; every rank must pass the same
    ``random_state`` seed / init.  The exchange steps are: global mean and variance, k-means++
    potentials and candidate rows, and ONE packed all-reduce of [sums | counts | stats] per Lloyd
    iteration.  ``labels_`` then covers the local rows; centres and inertia are global.
    """

    def __init__(self, n_clusters=8, *, init="k-means++", n_init="auto", max_iter=300, tol=1e-4, verbose=0,
                 random_state=None, copy_x=True, algorithm="lloyd", device=None, sharded=False,
                 process_group=None, _backend=None):
        self.n_clusters = n_clusters
        self.init = init
        self.n_init = n_init
        self.max_iter = max_iter
        self.tol = tol
        self.verbose = verbose
        self.random_state = random_state
        self.copy_x = copy_x
        self.algorithm = algorithm
        self.device = device
        self.sharded = sharded or process_group is not None
        self.process_group = process_group
        self._backend = _backend            # test hook: a CPU stand-in for _Device (gloo tests)
        self.batched = True                 # eight Lloyd iterations per host sync (dic_kmeans_lloyd_run); False = one

    # ---- helpers ---------------------------------------------------------------------------
    def _device(self):
        if self._backend is not None:
            return torch.device("cpu")
        if not torch.cuda.is_available():
            raise RuntimeError("KMeansB200 needs a CUDA device: there is no CPU fallback")
        return torch.device(self.device) if self.device is not None else torch.device("cuda", torch.cuda.current_device())

    def _to_device(self, X):
        dev = self._device()
        if isinstance(X, torch.Tensor):
            if X.dtype not in _DT:
                X = X.to(torch.float64)
            X = X.to(dev)
        else:
            X = np.asarray(X)
            if X.dtype not in (np.float32, np.float64):
                X = X.astype(np.float64)              # sklearn validate_data(dtype=[float64, float32])
            X = torch.from_numpy(np.ascontiguousarray(X)).to(dev)
        if X.dim() != 2:
            raise ValueError(f"Expected 2D array, got {X.dim()}D array instead")
        return X.contiguous()

    def _state(self, X, K):
        return (self._backend or _Device)(X, K)

    def _n_init(self):
        if self.n_init == "auto":
            return 1 if (isinstance(self.init, str) and self.init == "k-means++") or not isinstance(self.init, str) else 10
        return int(self.n_init)

    def _rows(self, st, comm, idx_global, offset):
        """Rows of the GLOBAL matrix by global index: owners contribute, the rest add zeros."""
        idx = idx_global - offset
        mine = (idx >= 0) & (idx < st.N)
        rows = torch.zeros((idx.numel(), st.D), dtype=st.X.dtype, device=st.dev)
        if bool(mine.any()) or not comm.on:
            rows[mine] = st.X[idx[mine]]
        comm.sum_(rows)
        return rows

    def _kmeans_plusplus(self, st, rng, comm, n_global, offset):
        """sklearn/cluster/_kmeans.py:224-281 on device; the RandomState is consumed in the same
        order and quantity as sklearn does (choice, then uniform(size=trials) per centre)."""
        K, N = self.n_clusters, st.N
        trials = 2 + int(np.log(K))
        centers = torch.empty((K, st.D), dtype=st.X.dtype, device=st.dev)
        # sklearn: random_state.choice(n, p=uniform) = one random_sample() through a uniform cdf
        first = min(int(rng.random_sample() * n_global), n_global - 1)
        centers[0] = self._rows(st, comm, torch.tensor([first], device=st.dev), offset)[0]
        closest = torch.empty(N, dtype=st.X.dtype, device=st.dev)
        pot = st.min_d2(centers[0:1], None, closest)
        comm.sum_(pot)
        pot = pot[0]
        # sklearn draws random_state.uniform(size=trials) once per centre; the stream does not depend on the
        # data, so the (K-1) x trials numbers are drawn here in the same order and uploaded ONCE
        draws = torch.from_numpy(np.stack([rng.uniform(size=trials) for _ in range(1, K)])
                                 if K > 1 else np.zeros((0, trials))).to(st.dev)
        for c in range(1, K):
            rv = draws[c - 1] * pot
            cum = torch.cumsum(closest.to(torch.float64), 0)
            if comm.on:      # global cumulative sum = local one + the totals of the lower ranks
                totals = comm.gather(cum[-1:] if N else cum.new_zeros(1)).flatten()
                base = totals[:comm.rank].sum()
                below = torch.cumsum(totals, 0)
                owner = torch.searchsorted(below, rv).clamp_(max=comm.size - 1)
                local = torch.searchsorted(cum, rv - base).clamp_(max=max(N - 1, 0))
                cand = torch.where(owner == comm.rank, local + offset, torch.full_like(local, -1))
                cand_rows = self._rows(st, comm, cand, offset) if N else self._rows(st, comm, cand, offset)
            else:
                cand = torch.searchsorted(cum, rv).clamp_(max=N - 1)
                cand_rows = st.X[cand].contiguous()
            pots = st.min_d2(cand_rows, closest, None)
            comm.sum_(pots)
            best = torch.argmin(pots)
            centers[c] = cand_rows[best]
            pot = st.min_d2(centers[c:c + 1], closest, closest)
            comm.sum_(pot)
            pot = pot[0]
        return centers

    def _relocate_empty(self, st, centers_old, comm):
        """sklearn _k_means_common.pyx:167-212 (rare path; torch ops on device).  Sharded: every
        rank offers its n_empty farthest rows, all ranks pick the same global winners and apply
        the same corrections to the (already global) sums and counts."""
        empty = torch.nonzero(st.counts == 0).flatten()
        n_empty = int(empty.numel())
        if n_empty == 0:
            return
        lab = st.labels.long()
        dist2 = ((st.X - centers_old[lab]) ** 2).sum(1)
        k = min(n_empty, st.N)
        top = torch.topk(dist2, k)
        cand_d = torch.full((n_empty,), -1.0, dtype=torch.float64, device=st.dev)
        cand_lab = torch.zeros(n_empty, dtype=torch.float64, device=st.dev)
        cand_row = torch.zeros((n_empty, st.D), dtype=torch.float64, device=st.dev)
        cand_d[:k] = top.values.to(torch.float64)
        cand_lab[:k] = lab[top.indices].to(torch.float64)
        cand_row[:k] = st.X[top.indices].to(torch.float64)
        all_d = comm.gather(cand_d).flatten()
        all_lab = comm.gather(cand_lab).flatten()
        all_row = comm.gather(cand_row).reshape(-1, st.D)
        if float(all_d.max()) <= 0.0:
            return
        order = torch.argsort(all_d, descending=True, stable=True)[:n_empty]   # farthest first
        for new_id, j in zip(empty.tolist(), order.tolist()):
            old_id = int(all_lab[j])
            row = all_row[j]
            st.sums[old_id] -= row
            st.sums[new_id] = row
            st.counts[new_id] = 1
            st.counts[old_id] -= 1

    LLOYD_BATCH = 8      # iterations enqueued per host synchronisation

    def _lloyd_batched(self, st, centers, tol_eff, comm):
        """_lloyd with the iterations enqueued LLOYD_BATCH at a time and the stopping rule evaluated on the device
        (dic_kmeans_lloyd_run): the kernels of a batch that follow the stopping iteration return at once, so labels,
        sums and centres are exactly those of the per-iteration loop.  An empty cluster (rare) stops the batch with
        the centres untouched; relocation and that iteration's M-step then run on the host path and the run resumes."""
        st.labels.fill_(-1)
        if st.status8 is not None:
            st.status8.zero_()
        strict = False
        it = 0
        while it < self.max_iter:
            status = st.lloyd_run(centers, DIC_KM_COUNT_CHANGES | DIC_KM_NO_INERTIA, min(self.LLOYD_BATCH, self.max_iter - it),
                                  tol_eff)
            it = int(status[5])
            flag = int(status[4])
            if flag == 1:
                strict = status[6] == 1.0
                break
            if flag == 2:                                        # the iteration `it` left an empty cluster
                changed = float(st.stats[1])
                self._relocate_empty(st, centers, comm)
                cnt = st.counts.clamp(min=1.0)[:, None]
                new = torch.where(st.counts[:, None] > 0, st.sums / cnt, centers.to(torch.float64)).to(centers.dtype)
                shift2 = float(((new - centers).to(torch.float64) ** 2).sum())
                centers.copy_(new)                               # in place: the cached kernel arguments stay valid
                if int(changed) == 0:
                    strict = True
                    break
                if shift2 <= tol_eff:
                    break
                st.status8[4:5].zero_()                         # resume; [5] keeps counting
        n_iter = it
        if strict:
            st.assign(centers, DIC_KM_KEEP_LABELS, want_sums=False)
        else:
            st.assign(centers, 0, want_sums=False)
        inertia = float(st.stats[0])
        return st.labels.clone(), inertia, centers, n_iter

    def _lloyd(self, st, centers, tol_eff, comm):
        """One run of sklearn/cluster/_kmeans.py:630-757."""
        if not comm.on and hasattr(st, "lloyd_run") and self.batched:
            return self._lloyd_batched(st, centers, tol_eff, comm)
        st.labels.fill_(-1)
        strict = False
        n_iter = 0
        for i in range(self.max_iter):
            n_iter = i + 1
            if not comm.on and hasattr(st, "lloyd_step"):
                changed, shift2, n_empty, _ = st.lloyd_step(centers, DIC_KM_COUNT_CHANGES | DIC_KM_NO_INERTIA)
                if n_empty > 0:                  # rare: the centres were left untouched, redo the M-step on the host
                    changed = None
            else:
                st.assign(centers, DIC_KM_COUNT_CHANGES | DIC_KM_NO_INERTIA)
                comm.sum_(st.sums, st.counts, st.stats)          # the one exchange step per iteration
                if hasattr(st, "update"):
                    changed, shift2, n_empty, _ = st.update(centers)    # in place, one host sync
                    if n_empty > 0:
                        changed = None
                else:
                    changed = None
            if changed is None:
                changed = float(st.stats[1])
                if int((st.counts == 0).sum()) > 0:
                    self._relocate_empty(st, centers, comm)
                cnt = st.counts.clamp(min=1.0)[:, None]
                new = torch.where(st.counts[:, None] > 0, st.sums / cnt, centers.to(torch.float64)).to(centers.dtype)
                shift2 = float(((new - centers).to(torch.float64) ** 2).sum())
                centers = new
            if int(changed) == 0:
                strict = True
                break
            if shift2 <= tol_eff:
                break
        if strict:
            st.assign(centers, DIC_KM_KEEP_LABELS, want_sums=False)
        else:
            st.assign(centers, 0, want_sums=False)
        comm.sum_(st.stats)
        inertia = float(st.stats[0])
        return st.labels.clone(), inertia, centers, n_iter

    def _same_as_best(self, labels, best_labels, K, comm):
        # K x K contingency table on the device (every label of one run maps to at most one label of the other:
        # _same_clustering); the labels stay where they are.  Restarts that find the same partition with permuted
        # labels differ in the last bits of their inertia (the order of the float32 run sums follows the labels), so
        # this test runs for many restarts: two D2H copies of the labels and a host-side np.unique each time cost
        # more than the restart's Lloyd iterations.
        table = torch.bincount(labels.long() * K + best_labels.long(), minlength=K * K)[:K * K].to(torch.float64)
        comm.sum_(table)
        return bool(((table.view(K, K) > 0).sum(1) <= 1).all())

    # ---- sklearn API -------------------------------------------------------------------------
    def fit(self, X, y=None, sample_weight=None):
        if sample_weight is not None:
            raise NotImplementedError("sample_weight is not supported (the reference never passes it)")
        as_tensor = isinstance(X, torch.Tensor)
        X = self._to_device(X)
        K = int(self.n_clusters)
        comm = _Comm(self.process_group if self.sharded else None) if self.sharded else _Comm.__new__(_Comm)
        if not self.sharded:
            comm.group, comm.on, comm.rank, comm.size = None, False, 0, 1
        n_local = torch.tensor([float(X.shape[0])], dtype=torch.float64, device=X.device)
        counts_all = comm.gather(n_local).flatten()
        n_global = int(counts_all.sum())
        offset = int(counts_all[:comm.rank].sum())
        if n_global < K:
            raise ValueError(f"n_samples={n_global} should be >= n_clusters={K}.")
        rng = _check_random_state(self.random_state)
        init = self.init
        init_arr = None
        if not isinstance(init, str):
            init_arr = torch.as_tensor(np.asarray(init), dtype=X.dtype).to(X.device)
            if tuple(init_arr.shape) != (K, X.shape[1]):
                raise ValueError(f"The shape of the initial centers {tuple(init_arr.shape)} does not match "
                                 f"the number of clusters {K} / features {X.shape[1]}.")
        elif init != "k-means++":
            raise NotImplementedError("init must be 'k-means++' or an array of centres")
        if comm.on:
            tot = X.sum(dim=0, dtype=torch.float64)
            comm.sum_(tot)
            mean = (tot / n_global).to(X.dtype)
        else:
            mean = X.mean(dim=0)
        Xc = X - mean                                                    # _kmeans.py:1487-1493
        if comm.on:
            ss = (Xc.to(torch.float64) ** 2).sum(0)
            s1 = Xc.sum(0, dtype=torch.float64)
            comm.sum_(ss, s1)
            var = ss / n_global - (s1 / n_global) ** 2
            tol_eff = float(var.mean()) * self.tol
        else:
            tol_eff = float(torch.var(Xc, dim=0, unbiased=False).mean()) * self.tol   # _kmeans.py:285-293
        st = self._state(Xc, K)
        best = None
        for _ in range(self._n_init()):
            c0 = (init_arr - mean) if init_arr is not None else self._kmeans_plusplus(st, rng, comm, n_global, offset)
            labels, inertia, centers, n_iter = self._lloyd(st, c0.contiguous(), tol_eff, comm)
            if best is None or (inertia < best[1] and not self._same_as_best(labels, best[0], K, comm)):
                best = (labels, inertia, centers, n_iter)
        labels, inertia, centers, n_iter = best
        centers = centers + mean                                         # _kmeans.py:1543-1546
        self._centers_dev = centers
        self.cluster_centers_ = centers if as_tensor else centers.cpu().numpy()
        self.labels_ = labels if as_tensor else labels.cpu().numpy()
        self.inertia_ = inertia
        self.n_iter_ = n_iter
        self.n_features_in_ = X.shape[1]
        return self

    def fit_predict(self, X, y=None, sample_weight=None):
        return self.fit(X, sample_weight=sample_weight).labels_

    def _centers_for(self, X):
        c = self.cluster_centers_
        c = c if isinstance(c, torch.Tensor) else torch.as_tensor(np.asarray(c))
        return c.to(device=X.device, dtype=X.dtype).contiguous()

    def predict(self, X):
        if not hasattr(self, "cluster_centers_"):
            raise RuntimeError("This KMeansB200 instance is not fitted yet")
        as_tensor = isinstance(X, torch.Tensor)
        X = self._to_device(X)
        centers = self._centers_for(X)
        st = self._state(X, centers.shape[0])
        st.assign(centers, 0, want_sums=False)
        return st.labels if as_tensor else st.labels.cpu().numpy()

    def score_distortion(self, X):
        """sum_i min_j ||x_i - c_j|| / N - the elbow distortion of p2_clustering_optK.py:261-264
        (local rows only when sharded; all-reduce numerator and N yourself)."""
        X = self._to_device(X)
        centers = self._centers_for(X)
        st = self._state(X, centers.shape[0])
        st.assign(centers, 0, want_sums=False)
        return float(st.stats[2]) / X.shape[0]
