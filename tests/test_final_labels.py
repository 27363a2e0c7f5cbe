"""p4 final labelling: the alignment helpers against the reference's own numpy code
(CPU, tiny), the whole k-means branch on the GPU against sklearn driven through the same steps."""
import numpy as np
import pytest
import torch


def _ref_generate_align_map(org_label, ob, padding):
    """Restatement of p4_clustering_final.py:63-98 in plain numpy (test-side checker)."""
    org_label = org_label.copy()
    v = ob[:, 0, :] * padding[:, 0, :]
    avg = v.sum(1) / padding[:, 0, :].sum(1)
    n_clusters = len(set(org_label)) - (1 if -1 in org_label else 0)
    sbp = [np.average(avg[org_label == i]) for i in range(n_clusters)]
    idx = [np.where(org_label == i) for i in range(n_clusters)]
    order = np.argsort(sbp)[::-1]
    amap = {prev: cur for cur, prev in enumerate(order)}
    amap = {k: amap[k] for k in sorted(amap)}
    for o, n in amap.items():
        org_label[idx[o]] = n
    return amap, org_label


def _cohort(n, K, D, seed):
    rng = np.random.RandomState(seed)
    centers = rng.normal(size=(K, D)) * 4
    z = rng.randint(0, K, size=n)
    hidden = (centers[z] + rng.normal(size=(n, D))).astype(np.float32)
    T, C = 12, 6
    pad = (rng.uniform(size=(n, C, T)) < 0.7).astype(np.float32)
    pad[:, :, 0] = 1
    ob = (rng.normal(size=(n, C, T)) + 100 + 10 * z[:, None, None]).astype(np.float32)
    return {"hidden": hidden, "ob": ob, "padding_mask": pad, "encounter_id": np.arange(n) + 1000 * seed}


@pytest.mark.gpu
def test_generate_align_map_matches_reference_numpy():
    from deep_interpolation_clustering_b200 import final_labels
    d = _cohort(500, 4, 8, 0)
    lab = np.random.RandomState(1).randint(0, 4, size=500)
    amap, aligned, cen = final_labels.generate_align_map(lab, d["ob"], d["padding_mask"], feat=d["hidden"])
    ramap, raligned = _ref_generate_align_map(lab, d["ob"], d["padding_mask"])
    assert amap == {int(k): int(v) for k, v in ramap.items()}
    assert np.array_equal(aligned, raligned)
    for i in range(4):
        assert np.allclose(cen[i], d["hidden"][raligned == i].astype(np.float64).mean(0), rtol=1e-6)
    assert np.array_equal(final_labels.align_labels(lab, amap), raligned)


@pytest.mark.gpu
def test_final_kmeans_labels_flow(tmp_path):
    """Fit on train, SBP-ordered alignment, centre permutation, predict on three cohorts, .npy contract."""
    from deep_interpolation_clustering_b200 import final_labels
    K, D = 4, 16
    tr, va, te = _cohort(3000, K, D, 0), _cohort(700, K, D, 1), _cohort(900, K, D, 2)
    km, amap, labels = final_labels.final_kmeans_labels(dict(tr), dict(va), dict(te), K, out_path=str(tmp_path),
                                                        n_init=3, random_state=0)
    # predict = nearest ALIGNED centre (float64 brute force as the checker)
    for cohort, data in zip(final_labels.COHORTS, (tr, va, te)):
        d2 = ((data["hidden"][:, None, :].astype(np.float64) - km.cluster_centers_[None].astype(np.float64)) ** 2).sum(2)
        assert np.array_equal(labels[cohort], d2.argmin(1))
        rec = final_labels.load_features(str(tmp_path / f"{cohort}_{K}.npy"))
        assert set(rec) == {"hidden", "encounter_id", "cluster_id"}             # ob / padding_mask dropped
        assert np.array_equal(rec["cluster_id"], labels[cohort])
    # aligned ids are ordered by descending mean SBP on train
    v = (tr["ob"][:, 0] * tr["padding_mask"][:, 0]).sum(1) / tr["padding_mask"][:, 0].sum(1)
    means = [v[labels["train"] == i].mean() for i in range(K)]
    assert all(means[i] > means[i + 1] for i in range(K - 1))
    assert sorted(amap.values()) == list(range(K))


def test_align_labels_keeps_noise_ids():
    from deep_interpolation_clustering_b200.final_labels import align_labels
    lab = np.array([0, 1, -1, 2, 1, 0])
    assert np.array_equal(align_labels(lab, {0: 2, 1: 0, 2: 1}), np.array([2, 0, -1, 1, 0, 2]))


@pytest.mark.gpu
def test_epoch_path_on_device():
    """latents accumulated in HBM, labels from the DEC kernel, label delta, k-means centre init."""
    from deep_interpolation_clustering_b200 import epoch_path
    from oracle import dec_oracle
    dev = torch.device("cuda:0")
    d = _cohort(2000, 4, 32, 5)
    feats = epoch_path.EpochFeatures(2000, 32, dev)
    for i in range(0, 2000, 256):
        feats.append(torch.from_numpy(d["hidden"][i:i + 256]).to(dev))
    assert feats.n == 2000 and torch.equal(feats.features().cpu(), torch.from_numpy(d["hidden"]))
    pred, centers, km = epoch_path.init_cluster_centers(feats.features(), 4, n_init=3, random_state=0)
    assert centers.requires_grad and centers.shape == (4, 32) and centers.dtype == torch.float32
    lab = feats.cluster_pred(centers)
    q = dec_oracle.soft_assign(d["hidden"].astype(np.float64), centers.detach().cpu().numpy().astype(np.float64), 1.0)
    assert np.array_equal(lab.cpu().numpy(), q.argmax(1))
    assert np.array_equal(lab.cpu().numpy(), np.asarray(pred.cpu() if isinstance(pred, torch.Tensor) else pred))
    assert epoch_path.label_delta(lab, None) == 1.0
    prev = lab.clone()
    prev[:100] = (prev[:100] + 1) % 4
    assert abs(epoch_path.label_delta(lab, prev) - 0.05) < 1e-12
    _, rc, _ = epoch_path.init_cluster_centers(feats.features(), 4, mode="random", rng=np.random.RandomState(0))
    lo, hi = d["hidden"].min(0), d["hidden"].max(0)
    rcn = rc.detach().cpu().numpy()
    assert (rcn >= lo - 1e-6).all() and (rcn <= hi + 1e-6).all()
