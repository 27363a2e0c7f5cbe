"""Cluster-validity metrics called from inside the gap loop (p2_clustering_optK.py:401-405).

Mirror of internal_eval.py:112-147 (same class names, ``metric(x, labels)`` call signature).
SURVEY.md section 8(f) ranks these as the first "next" row after the hot path.  Status:
Calinski-Harabasz and Davies-Bouldin are O(N K D) and run as device reductions.  Silhouette is
O(N^2): for D <= 256 its per-row, per-cluster distance sums come from the tcgen05 tile kernel
(``dic_cluster_rowsums``: rows sorted by cluster, clusters padded to whole 128-row tiles, nothing
materialised); other shapes and Dunn use chunked device distance tiles (torch.cdist, library code).
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


def _centroids(x, inv, K):
    sums = torch.zeros((K, x.shape[1]), dtype=x.dtype, device=x.device).index_add_(0, inv, x)
    cnt = torch.bincount(inv, minlength=K).to(x.dtype)
    return sums / cnt[:, None], cnt


class CHIndex(object):
    """Calinski-Harabasz: [tr(B)/(K-1)] / [tr(W)/(N-K)].  internal_eval.py:125-135."""

    def __call__(self, x, labels, *args, **kwargs):
        x, inv, K = _prep(x, labels)
        N = x.shape[0]
        cen, cnt = _centroids(x, inv, K)
        mean = x.mean(0)
        extra = float((cnt * ((cen - mean) ** 2).sum(1)).sum())
        intra = float(((x - cen[inv]) ** 2).sum())
        return 1.0 if intra == 0.0 else extra * (N - K) / (intra * (K - 1.0))


class DBIndex(object):
    """Davies-Bouldin: mean_i max_{j != i} (s_i + s_j) / d_ij.  internal_eval.py:138-147."""

    def __call__(self, x, label, *args, **kwargs):
        x, inv, K = _prep(x, label)
        cen, cnt = _centroids(x, inv, K)
        d = torch.sqrt(((x - cen[inv]) ** 2).sum(1))
        s = torch.zeros(K, dtype=x.dtype, device=x.device).index_add_(0, inv, d) / cnt
        cd = torch.cdist(cen, cen)
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

    def __call__(self, x, labels, *args, **kwargs):
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
    """min nearest inter-cluster distance / max cluster diameter.  internal_eval.py:15-109."""

    def __init__(self, chunk=8192):
        self.chunk = chunk

    def __call__(self, x, labels, *args, **kwargs):
        x, inv, K = _prep(x, labels)
        N = x.shape[0]
        min_inter = torch.tensor(float("inf"), dtype=x.dtype, device=x.device)
        max_diam = torch.zeros((), dtype=x.dtype, device=x.device)
        for i0 in range(0, N, self.chunk):
            d = torch.cdist(x[i0:i0 + self.chunk], x)
            same = inv[i0:i0 + self.chunk, None] == inv[None, :]
            max_diam = torch.maximum(max_diam, torch.where(same, d, torch.zeros_like(d)).max())
            inter = torch.where(same | (d == 0), torch.full_like(d, float("inf")), d)
            min_inter = torch.minimum(min_inter, inter.min())
        return float(min_inter / max_diam)
