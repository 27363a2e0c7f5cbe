"""Final cluster labelling with the chosen K.

Mirror of the k-means branch of ``Cluster.pred`` and its alignment helpers,
p4_clustering_final.py:63-98 (generate_align_map), :101-111 (align_labels), :141-179 (pred):

    KMeans(n_clusters=K, init='k-means++', n_init=20).fit(train hidden)      (:159)
    predict(train) -> order the clusters by DESCENDING mean of the first vital (SBP) -> align map
    permute cluster_centers_ with the map                                     (:164-166)
    predict(train / valid / test) with the aligned centres, save {..., 'cluster_id'} as .npy   (:168-179)

The k-means fit / predict run on the B200 (KMeansB200: the same Lloyd and k-means++ kernels as the
gap sweep); the per-encounter SBP means and per-cluster averages are device reductions; the on-disk
contract is the reference's: ``np.save`` of a dict with ``encounter_id, hidden[, cluster_pred,
cluster_label], cluster_id`` and WITHOUT ``ob`` / ``padding_mask`` (:175-176).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from .kmeans import KMeansB200

COHORTS = ("train", "valid", "test")         # p4_clustering_final.py:27


def _dev(a, device):
    if isinstance(a, torch.Tensor):
        return a.to(device)
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a))).to(device)


def generate_align_map(org_label, ob, padding, feat=None, device=None):
    """p4_clustering_final.py:63-98.  ``org_label`` (N,) raw cluster ids, ``ob`` / ``padding`` (N, C, T).
    Returns ``(align_map, aligned_label, new_feat_centers)`` like the reference: clusters are renumbered by
    descending average of each encounter's mean first-vital value (SBP) over its valid slots."""
    device = torch.device(device or "cuda")
    lab = _dev(np.asarray(org_label), device).long()
    v = _dev(ob, device)[:, 0, :].to(torch.float64)
    m = _dev(padding, device)[:, 0, :].to(torch.float64)
    avg = (v * m).sum(1) / m.sum(1)                                              # :77-79
    present = torch.unique(lab)
    n_clusters = int(present.numel()) - (1 if bool((present == -1).any()) else 0)   # :80
    valid = lab >= 0
    cnt = torch.bincount(lab[valid], minlength=n_clusters).to(torch.float64)
    tot = torch.zeros(n_clusters, dtype=torch.float64, device=device).index_add_(0, lab[valid], avg[valid])
    cluster_sbp = (tot / cnt).cpu().numpy()                                      # np.average per cluster, :84
    sorted_cluster_ids = np.argsort(cluster_sbp)[::-1]                           # :86
    align_map = {int(prev): cur for cur, prev in enumerate(sorted_cluster_ids)}
    align_map = {k: align_map[k] for k in sorted(align_map)}                     # :87-88
    aligned = align_labels(org_label, align_map)
    new_feat_centers = []
    if feat is not None:                                                         # :95-98
        f = _dev(feat, device).to(torch.float64)
        al = _dev(aligned, device).long()
        for i in range(n_clusters):
            new_feat_centers.append(f[al == i].mean(0).cpu().numpy())
    return align_map, aligned, new_feat_centers


def align_labels(org_label, align_map):
    """p4_clustering_final.py:101-111: relabel through the map (ids outside the map, e.g. -1, are kept)."""
    org = np.asarray(org_label)
    lut = np.arange(max(int(org.max()) + 1, max(align_map) + 1) if org.size else 1)
    for k, v in align_map.items():
        lut[k] = v
    out = org.copy()
    ok = org >= 0
    out[ok] = lut[org[ok]]
    return out


def final_kmeans_labels(train_data, valid_data, test_data, num_clusters, out_path=None, n_init=20, overwrite=False,
                        random_state=None, device=None):
    """The k-means branch of p4_clustering_final.py:141-179.  Each ``*_data`` is the reference's feature dict
    (``hidden`` (N, D), ``ob`` / ``padding_mask`` (N, C, T), ``encounter_id``, ...).  Returns
    ``(kmeans_model, align_map, {cohort: cluster_id})``; writes ``{cohort}_{K}.npy`` under ``out_path`` when given."""
    km = KMeansB200(n_clusters=num_clusters, init="k-means++", n_init=n_init, random_state=random_state,
                    device=device).fit(train_data["hidden"])                     # :159
    train_raw = km.predict(train_data["hidden"])                                 # :160
    align_map, _, _ = generate_align_map(np.asarray(train_raw), train_data["ob"], train_data["padding_mask"],
                                         device=device)                          # :161
    idp = np.array(km.cluster_centers_, copy=True)                               # :164-166
    for org_id, new_id in align_map.items():
        km.cluster_centers_[new_id] = idp[org_id]
    labels = {}
    if out_path:
        os.makedirs(out_path, exist_ok=True)
    for cohort, data in zip(COHORTS, (train_data, valid_data, test_data)):
        if data is None:
            continue
        cid = np.asarray(km.predict(data["hidden"]))                             # :173
        labels[cohort] = cid
        if out_path:
            f = os.path.join(out_path, f"{cohort}_{num_clusters}.npy")
            if os.path.exists(f) and not overwrite:                              # :169-171
                continue
            rec = {k: v for k, v in data.items() if k not in ("ob", "padding_mask")}   # :175-176
            rec["cluster_id"] = cid
            np.save(f, rec)
    return km, align_map, labels


def load_features(path):
    """Reader of the reference's feature files (``np.save`` of a dict; pretrain_trainer.py:101-113,
    p2_clustering_optK.py:51-62): ``np.load(..., allow_pickle=True).item()``."""
    return np.load(path, allow_pickle=True).item()
