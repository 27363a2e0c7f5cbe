"""Cluster-validity metrics called from inside the gap loop (p2_clustering_optK.py:401-405).

Mirror of internal_eval.py:112-147 (same class names, ``metric(x, labels)`` call signature).
SURVEY.md section 8(f) ranks these as the first "next" row after the hot path.  Status:
Calinski-Harabasz and Davies-Bouldin are two passes over X on the device (the Lloyd pass in its keep-labels form for
the centroids + ``dic_cluster_scatter`` for the dispersions), exact also when the rows are sharded over ranks
(``group=``).  Silhouette is O(N^2): for D <= 256 its per-row, per-cluster distance sums come from the tcgen05 tile
kernel (``dic_cluster_rowsums``: rows sorted by cluster, clusters padded to whole 128-row tiles, nothing
materialised; other shapes use chunked torch.cdist tiles, library code).  Dunn's two extrema over all pairs come from
``dic_dunn_minmax`` (register tiles, upper triangle).
Formulas follow sklearn.metrics 1.9.0 (_unsupervised.py).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

TC_MIN_ROWS = 1024     # below this the chunked float64 path is used


def _prep(x, labels):
    if not isinstance(x, torch.Tensor):
        x = torch.from_numpy(np.ascontiguousarray(np.asarray(x))).cuda()
    lab = torch.as_tensor(np.asarray(labels) if not isinstance(labels, torch.Tensor) else labels).to(x.device).long()
    uniq, inv = torch.unique(lab, return_inverse=True)
    return x.to(torch.float64), inv, int(uniq.numel())


_DT = {torch.float32: 0, torch.float64: 1}


def _prep_native(x, labels, group=None):
    """X in its own dtype (float32 / float64, contiguous, on the device), labels as int32 ranks 0..K-1 of the
    distinct label values (over ALL ranks when the rows are sharded)."""
    if not isinstance(x, torch.Tensor):
        x = torch.from_numpy(np.ascontiguousarray(np.asarray(x))).cuda()
    if x.dtype not in _DT:
        x = x.to(torch.float64)
    x = x.contiguous()
    lab = torch.as_tensor(np.asarray(labels) if not isinstance(labels, torch.Tensor) else labels).to(x.device).long()
    if _sharded(group):
        import torch.distributed as dist
        hi = lab.max().reshape(1) if lab.numel() else torch.zeros(1, dtype=torch.long, device=x.device)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
        present = torch.zeros(int(hi) + 1, dtype=torch.long, device=x.device)
        present[lab] = 1
        dist.all_reduce(present, op=dist.ReduceOp.MAX, group=group)
        rank_of = torch.cumsum(present, 0) - 1
        return x, rank_of[lab].to(torch.int32).contiguous(), int(present.sum())
    uniq, inv = torch.unique(lab, return_inverse=True)
    return x, inv.to(torch.int32).contiguous(), int(uniq.numel())


def _sharded(group):
    import torch.distributed as dist
    return group is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def cluster_stats(x, lab32, K, group=None):
    """Per-cluster counts, centroids and dispersion of rows `x` with labels 0..K-1, all on the device:
        counts (K), centroids (K, D), s1[k] = sum ||x - c_k||, s2[k] = sum ||x - c_k||^2, grand mean (D)   [float64]
    Two passes over X: the Lloyd pass in its keep-labels form (dic_kmeans_assign: sums / counts) and
    dic_cluster_scatter.  With a process group the rows are this rank's shard: sums / counts and the (K, 2)
    dispersion table are all-reduced (K D + 3 K doubles), so every rank returns the statistics of ALL rows."""
    L = _lib.lib()
    dev, (N, D) = x.device, x.shape
    if K > 64 or D > 512:
        raise ValueError(f"cluster_stats covers K <= 64 and D <= 512 (got K={K}, D={D})")
    dt = _DT[x.dtype]
    sums = torch.zeros((K, D), dtype=torch.float64, device=dev)
    counts = torch.zeros(K, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        st = _lib.current_stream(dev)
        if N:
            stats = torch.empty(4, dtype=torch.float64, device=dev)
            ws = torch.empty(max(int(L.dic_kmeans_workspace_bytes(max(K, 16), D)), 16), dtype=torch.uint8, device=dev)
            cen0 = torch.zeros((K, D), dtype=x.dtype, device=dev)
            _lib.check(L.dic_kmeans_assign(_lib.ptr(x), _lib.ptr(cen0), _lib.ptr(lab32), _lib.ptr(sums), _lib.ptr(counts),
                                           _lib.ptr(stats), _lib.ptr(ws), N, D, K, dt, 2 | 4, st), "dic_kmeans_assign")
        if _sharded(group):
            import torch.distributed as dist
            packed = torch.cat([sums.reshape(-1), counts])
            dist.all_reduce(packed, group=group)
            sums, counts = packed[:K * D].view(K, D), packed[K * D:]
        cen = sums / counts.clamp(min=1.0)[:, None]
        out = torch.zeros((K, 2), dtype=torch.float64, device=dev)
        if N:
            ws2 = torch.empty(max(int(L.dic_cluster_scatter_workspace_bytes(K)), 16), dtype=torch.uint8, device=dev)
            cen_x = cen.to(x.dtype).contiguous()
            _lib.check(L.dic_cluster_scatter(_lib.ptr(x), _lib.ptr(lab32), _lib.ptr(cen_x), _lib.ptr(out), _lib.ptr(ws2),
                                             N, D, K, dt, st), "dic_cluster_scatter")
        if _sharded(group):
            import torch.distributed as dist
            dist.all_reduce(out, group=group)
    mean = sums.sum(0) / counts.sum()
    return counts, cen, out[:, 0], out[:, 1], mean


class CHIndex(object):
    """Calinski-Harabasz: [tr(B)/(K-1)] / [tr(W)/(N-K)].  internal_eval.py:125-135
    (sklearn.metrics.calinski_harabasz_score).  One keep-labels Lloyd pass + one dispersion pass on the device;
    the rest is K-sized.  `group`: the rows are sharded over the ranks of a process group."""

    def __call__(self, x, labels, *args, group=None, **kwargs):
        x, lab, K = _prep_native(x, labels, group)
        cnt, cen, _, s2, mean = cluster_stats(x, lab, K, group)
        N = float(cnt.sum())
        extra = float((cnt * ((cen - mean) ** 2).sum(1)).sum())
        intra = float(s2.sum())
        return 1.0 if intra == 0.0 else extra * (N - K) / (intra * (K - 1.0))


class DBIndex(object):
    """Davies-Bouldin: mean_i max_{j != i} (s_i + s_j) / d_ij.  internal_eval.py:138-147
    (sklearn.metrics.davies_bouldin_score); same two device passes as CHIndex."""

    def __call__(self, x, label, *args, group=None, **kwargs):
        x, lab, K = _prep_native(x, label, group)
        cnt, cen, s1, _, _ = cluster_stats(x, lab, K, group)
        s = s1 / cnt
        cd = torch.sqrt(((cen[:, None, :] - cen[None, :, :]) ** 2).sum(2))
        if torch.allclose(s, torch.zeros_like(s)) or torch.allclose(cd, torch.zeros_like(cd)):
            return 0.0
        cd[cd == 0] = float("inf")
        scores = ((s[:, None] + s[None, :]) / cd).max(dim=1).values
        return float(scores.mean())


class Sihouette(object):
    """Mean silhouette coefficient (name as upstream).  internal_eval.py:112-122."""

    def __init__(self, chunk=8192, native=True):
        self.chunk = chunk
        self.native = native

    @staticmethod
    def _finish(per_cluster, own, cnt):
        """(rows, K) distance sums -> sum of silhouette samples (sklearn _unsupervised.py:silhouette_samples)."""
        rows = torch.arange(own.numel(), device=own.device)
        n_own = cnt[own]
        a = per_cluster[rows, own] / (n_own - 1).clamp(min=1)
        other = per_cluster / cnt[None, :]
        other[rows, own] = float("inf")
        b = other.min(dim=1).values
        sil = (b - a) / torch.maximum(a, b)
        sil = torch.where(n_own > 1, sil, torch.zeros_like(sil))         # singleton clusters score 0
        return torch.nan_to_num(sil).sum()

    @staticmethod
    def rowsums_native(x32, inv, K):
        """Per-row per-cluster distance sums (N, K) float64 in the ORIGINAL row order, via dic_cluster_rowsums."""
        dev = x32.device
        N, D = x32.shape
        cnt = torch.bincount(inv, minlength=K)
        order = torch.argsort(inv, stable=True)
        start = torch.cumsum(cnt, 0) - cnt                                # first sorted index of each cluster
        padded = (cnt + 127) // 128 * 128
        pstart = torch.cumsum(padded, 0) - padded                         # first packed row of each cluster
        n_pad = int(padded.sum())
        k_sorted = inv[order]
        pos = pstart[k_sorted] + (torch.arange(N, device=dev) - start[k_sorted])
        perm = torch.full((n_pad,), -1, dtype=torch.int32, device=dev)
        perm[pos] = order.to(torch.int32)
        tile_cluster = torch.repeat_interleave(torch.arange(K, device=dev, dtype=torch.int32), padded // 128)
        xc = (x32 - x32.mean(dim=0, keepdim=True)).contiguous()           # distances are translation invariant
        rowsum = torch.empty((n_pad, K), dtype=torch.float64, device=dev)
        L = _lib.lib()
        ws = torch.empty(int(L.dic_pairwise_workspace_bytes(n_pad, D)), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.dic_cluster_rowsums(_lib.ptr(xc), _lib.ptr(perm), _lib.ptr(tile_cluster), _lib.ptr(rowsum),
                                             _lib.ptr(ws), n_pad, D, K, _lib.current_stream(dev)),
                       "dic_cluster_rowsums")
        out = torch.empty((N, K), dtype=torch.float64, device=dev)
        out[order] = rowsum[pos]
        if bool(torch.isnan(rowsum[0, 0])):
            raise _lib.DicError("dic_cluster_rowsums: the tensor-core pipeline reported a timeout")
        return out

    def __call__(self, x, labels, *args, group=None, **kwargs):
        if _sharded(group):      # O(N^2) over ALL rows: gather the shards, every rank evaluates the same score
            x = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(x))).cuda()
            labels = torch.as_tensor(np.asarray(labels) if not isinstance(labels, torch.Tensor) else labels).to(x.device)
            x, labels = _all_gather_rows(x.contiguous(), group), _all_gather_rows(labels.contiguous(), group)
        x, inv, K = _prep(x, labels)
        N = x.shape[0]
        cnt = torch.bincount(inv, minlength=K).to(x.dtype)
        if self.native and x.is_cuda and N >= TC_MIN_ROWS and x.shape[1] <= 256 and x.shape[1] % 4 == 0:
            per_cluster = self.rowsums_native(x.to(torch.float32), inv, K)
            return float(self._finish(per_cluster, inv, cnt) / N)
        onehot = torch.zeros((N, K), dtype=x.dtype, device=x.device)
        onehot[torch.arange(N, device=x.device), inv] = 1.0
        total = torch.zeros((), dtype=x.dtype, device=x.device)
        for i0 in range(0, N, self.chunk):
            d = torch.cdist(x[i0:i0 + self.chunk], x)                    # (c, N) distances
            per_cluster = d @ onehot                                     # (c, K) distance sums
            own = inv[i0:i0 + self.chunk]
            rows = torch.arange(own.numel(), device=x.device)
            n_own = cnt[own]
            a = per_cluster[rows, own] / (n_own - 1).clamp(min=1)
            other = per_cluster / cnt[None, :]
            other[rows, own] = float("inf")
            b = other.min(dim=1).values
            sil = (b - a) / torch.maximum(a, b)
            sil = torch.where(n_own > 1, sil, torch.zeros_like(sil))     # singleton clusters score 0
            total += torch.nan_to_num(sil).sum()
        return float(total / N)


class DunnIndex(object):
    """min nearest inter-cluster distance / max cluster diameter.  internal_eval.py:15-109 ("nearest" inter-cluster
    distances, "farthest" diameters, cluster pairs at distance zero dropped like ``ic_distances.nonzero()``).  The
    (K, K) nearest-distance table and the largest diameter come from dic_dunn_minmax (64 x 64 register tiles, upper triangle); the reference's n x n matrix and
    its pure-Python double loop are gone.  Row-sharded input is all-gathered first (the metric is O(N^2) over ALL
    rows whichever rank holds them)."""

    def __call__(self, x, labels, *args, group=None, **kwargs):
        x, lab, K = _prep_native(x, labels, group)
        if _sharded(group):
            x, lab = _all_gather_rows(x, group), _all_gather_rows(lab, group)
        out = torch.empty(K * K + 1, dtype=torch.float64, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().dic_dunn_minmax(_lib.ptr(x), _lib.ptr(lab), _lib.ptr(out), x.shape[0], x.shape[1], K,
                                                  _DT[x.dtype], _lib.current_stream(x.device)), "dic_dunn_minmax")
        out = out.cpu().numpy()
        ic = out[:K * K].reshape(K, K)[np.triu_indices(K, 1)]          # nearest distance of every cluster pair
        ic = ic[ic != 0]       # :106 `ic_distances[ic_distances.nonzero()]`: a pair of clusters that touch drops out
        if ic.size == 0:
            raise ValueError("min() iterable argument is empty")       # what the reference raises in that case
        return float(ic.min() / out[K * K])


def _all_gather_rows(t, group):
    """Concatenation of every rank's rows (ragged shards: sizes are exchanged first)."""
    import torch.distributed as dist
    ws = dist.get_world_size(group)
    n = torch.tensor([t.shape[0]], dtype=torch.long, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(ws)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(v) for v in sizes]
    pad = max(sizes)
    buf = torch.zeros((pad,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    buf[:t.shape[0]] = t
    parts = [torch.empty_like(buf) for _ in range(ws)]
    dist.all_gather(parts, buf, group=group)
    return torch.cat([p[:m] for p, m in zip(parts, sizes)]).contiguous()
