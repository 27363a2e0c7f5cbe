"""Epoch-level feature / label path kept on the device.

The reference moves every batch to the host, deep-copies it, extends Python lists and re-stacks them
once per epoch (clustering_trainer.py:409-416 ``eval_one_epoch``, :486-493 ``merge_ob_pred``) only to
take an argmax and a label delta (:473-484 ``generate_pred_cluster``) or to seed the cluster centres
with k-means (:72-82).  Here the per-batch latents stay in HBM and the consumers are the kernels of the
hot path:

    EpochFeatures.append(hidden[, q])        <- eval loop, no D2H
    EpochFeatures.cluster_pred(centres)      -> hard labels from dic_dec_q_fwd (argmax_j q, :476)
    label_delta(pred, prev)                  -> fraction of changed labels (:477-483), ONE scalar to the host
    init_cluster_centers(hidden, K, ...)     -> KMeansB200(n_init=20).fit_predict + centres tensor (:72-79),
                                                or the uniform 'random' init (:83-90)
"""
from __future__ import annotations

import numpy as np
import torch

from . import functional as F_
from .kmeans import KMeansB200


class EpochFeatures:
    """Device-side replacement for the ob_pred_lst / merge_ob_pred pair (clustering_trainer.py:409-416,486-493)
    restricted to what the clustering loop consumes: the latent ``hidden`` and, optionally, the soft
    assignment of each batch.  Preallocates (capacity, D) once; ``append`` is a device copy.
    ``capacity=None``: grows by doubling (a data loader without ``len``)."""

    def __init__(self, capacity, dim, device, n_clusters=None):
        self.growable = capacity is None
        capacity = 4096 if capacity is None else capacity
        self.hidden = torch.empty((capacity, dim), dtype=torch.float32, device=device)
        self.q = torch.empty((capacity, n_clusters), dtype=torch.float32, device=device) if n_clusters else None
        self.n = 0

    def _grow(self, need):
        cap = max(2 * self.hidden.shape[0], need)
        for name in ("hidden", "q"):
            old = getattr(self, name)
            if old is not None:
                new = torch.empty((cap, old.shape[1]), dtype=old.dtype, device=old.device)
                new[:self.n].copy_(old[:self.n])
                setattr(self, name, new)

    def append(self, hidden, q=None):
        b = hidden.shape[0]
        if self.n + b > self.hidden.shape[0]:
            if not self.growable:
                raise ValueError(f"EpochFeatures capacity {self.hidden.shape[0]} exceeded")
            self._grow(self.n + b)
        self.hidden[self.n:self.n + b].copy_(hidden.detach(), non_blocking=True)
        if q is not None and self.q is not None:
            self.q[self.n:self.n + b].copy_(q.detach(), non_blocking=True)
        self.n += b

    def features(self):
        return self.hidden[:self.n]

    def cluster_pred(self, cluster_centers=None, alpha=1.0):
        """np.argmax(cluster_pred, axis=1) of clustering_trainer.py:476 as int32 device labels: from the stored
        q when there is one, otherwise recomputed from the latents in one dic_dec_q_fwd pass."""
        if self.q is not None and cluster_centers is None:
            return torch.argmax(self.q[:self.n], dim=1).to(torch.int32)
        return F_.dec_assign(self.features(), cluster_centers.detach(), alpha)["labels"]

    def reset(self):
        self.n = 0


def label_delta(cluster_pred, prev_pred):
    """clustering_trainer.py:477-483: 1.0 without a previous prediction, else the fraction of changed labels."""
    if prev_pred is None:
        return 1.0
    a = cluster_pred if isinstance(cluster_pred, torch.Tensor) else torch.as_tensor(np.asarray(cluster_pred))
    b = prev_pred if isinstance(prev_pred, torch.Tensor) else torch.as_tensor(np.asarray(prev_pred))
    b = b.to(a.device)
    return float((a.long() != b.long()).sum()) / b.shape[0]


def init_cluster_centers(hidden, cluster_number, mode="kmeans", n_init=20, random_state=None, rng=None):
    """clustering_trainer.py:72-90.  ``mode='kmeans'``: KMeans(n_clusters, n_init=20).fit_predict(hidden) ->
    (labels, centres as a float32 tensor with requires_grad, the fitted estimator for predict on the validation
    latents, :81-82).  ``mode='random'``: uniform in the per-dimension range of ``hidden`` (:85-90)."""
    x = hidden if isinstance(hidden, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(hidden)).cuda()
    if mode == "kmeans":
        km = KMeansB200(n_clusters=cluster_number, n_init=n_init, random_state=random_state)
        pred = km.fit_predict(x)
        cc = km.cluster_centers_                 # tensor when fitted on a tensor, ndarray otherwise (sklearn contract)
        cc = cc.detach() if isinstance(cc, torch.Tensor) else torch.from_numpy(np.asarray(cc))
        centers = cc.to(device=x.device, dtype=torch.float).clone().requires_grad_(True)
        return pred, centers, km
    if mode == "random":
        rng = np.random if rng is None else rng
        hi, lo = x.max(dim=0).values.cpu().numpy(), x.min(dim=0).values.cpu().numpy()
        c = rng.uniform(low=lo, high=hi, size=(cluster_number, hi.shape[-1]))
        return None, torch.tensor(c, dtype=torch.float, device=x.device, requires_grad=True), None
    raise NotImplementedError(mode)
