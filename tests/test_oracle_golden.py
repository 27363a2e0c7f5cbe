"""Pin the CPU oracle against fixtures produced by the reference itself.

Fixtures: tests/golden/*.npz written by oracle/gen_golden.py (reference run in
float32 and float64).  float64 restatement vs float64 reference must agree to
rounding (1e-10); the float32 reference is within its own rounding of that truth.
"""
import numpy as np
import pytest

from oracle import dec_oracle, interp_oracle, kmeans_oracle

INTERP_CASES = ["interp_c1", "interp_c2", "interp_c5", "interp_smoke", "interp_odd", "interp_dense", "interp_kernels"]
DEC_CASES = ["dec_k4", "dec_k16", "dec_alpha2", "dec_k10"]


def _close(a, b, rtol, atol, what):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    err = np.abs(a - b)
    lim = atol + rtol * np.abs(b)
    assert np.all(err <= lim), f"{what}: max err {err.max():.3e}, worst ratio {(err / lim).max():.2f}"


@pytest.mark.parametrize("case", INTERP_CASES)
def test_interp_forward_f64(golden, case):
    g = golden(case)
    C = g["x"].shape[1] // 4
    x, rt = g["x"].astype(np.float64), g["ref_t"].astype(np.float64)
    s = interp_oracle.sci_forward(x, g["sci_kernel"].astype(np.float64), rt, C)
    _close(s, g["sci_out_f64"], 1e-10, 1e-12, "sci")
    c = interp_oracle.cci_forward(g["sci_out_f64"], g["cci_kernel"].astype(np.float64), C)
    _close(c, g["cci_out_f64"], 1e-10, 1e-12, "cci")
    r = interp_oracle.rbf_forward(g["v"].astype(np.float64), x, g["rbf_kernel"].astype(np.float64), rt, C)
    _close(r, g["rbf_out_f64"], 1e-10, 1e-12, "rbf")


@pytest.mark.parametrize("case", INTERP_CASES)
def test_interp_backward_f64(golden, case):
    g = golden(case)
    C = g["x"].shape[1] // 4
    x, rt = g["x"].astype(np.float64), g["ref_t"].astype(np.float64)
    du, dK = interp_oracle.cci_backward(g["sci_out_f64"], g["cci_kernel"].astype(np.float64), C,
                                        g["g_cci"].astype(np.float64))
    _close(dK, g["d_cci_kernel_f64"], 1e-9, 1e-10, "d cci.kernel")
    _close(du, g["d_sci_out_f64"], 1e-9, 1e-10, "d sci_out")
    dk = interp_oracle.sci_backward(x, g["sci_kernel"].astype(np.float64), rt, C, g["d_sci_out_f64"])
    _close(dk, g["d_sci_kernel_f64"], 1e-9, 1e-10, "d sci.kernel")
    dv, dkr = interp_oracle.rbf_backward(g["v"].astype(np.float64), x, g["rbf_kernel"].astype(np.float64),
                                         rt, C, g["g_rbf"].astype(np.float64))
    _close(dv, g["dv_f64"], 1e-9, 1e-10, "dv")
    _close(dkr, g["d_rbf_kernel_f64"], 1e-9, 1e-10, "d rbf.kernel")


@pytest.mark.parametrize("case", INTERP_CASES)
def test_interp_f32_reference_is_near_f64_truth(golden, case):
    """Documents the reference's own float32 error, which bounds any 'rtol 1e-5' claim."""
    g = golden(case)
    for k in ("sci_out", "cci_out", "rbf_out"):
        _close(g[k], g[k + "_f64"], 1e-4, 1e-4, k)


def test_interp_allmasked_semantics(golden):
    """All-masked channel: w = -inf, y = y' = NaN in the reference; oracle replicates."""
    g = golden("interp_allmasked")
    C = 6
    s = interp_oracle.sci_forward(g["x"].astype(np.float64), g["sci_kernel"].astype(np.float64),
                                  g["ref_t"].astype(np.float64), C)
    ref = g["sci_out_f64"]
    assert np.array_equal(np.isnan(s), np.isnan(ref))
    assert np.array_equal(np.isneginf(s), np.isneginf(ref))
    ok = np.isfinite(ref)
    _close(s[ok], ref[ok], 1e-10, 1e-12, "sci finite part")
    assert np.isnan(ref[1, :, 2]).all() and np.isneginf(ref[1, :, 6 + 2]).all()


def test_linspace_grid_matches_torch(golden):
    for case in INTERP_CASES:
        g = golden(case)
        rt = interp_oracle.linspace_grid(float(g["hours"]), int(g["R"]), np.float32)
        assert np.array_equal(rt, g["ref_t"]), case


@pytest.mark.parametrize("case", DEC_CASES)
def test_dec_f64(golden, case):
    g = golden(case)
    a = float(g["alpha"])
    z, mu = g["z"].astype(np.float64), g["mu"].astype(np.float64)
    q = dec_oracle.soft_assign(z, mu, a)
    _close(q, g["q_f64"], 1e-11, 1e-14, "q")
    p = dec_oracle.target_distribution(q)
    _close(p, g["p_f64"], 1e-11, 1e-14, "p")
    _close(dec_oracle.kl_div_batchmean(p, q), g["kl_f64"], 1e-10, 1e-14, "kl")
    dz, dmu = dec_oracle.kl_backward_closed_form(z, mu, p, a)
    _close(dz, g["dz_kl_f64"], 1e-9, 1e-13, "dz kl")
    _close(dmu, g["dmu_kl_f64"], 1e-9, 1e-13, "dmu kl")
    dz, dmu = dec_oracle.soft_assign_backward(z, mu, g["gq"].astype(np.float64), a)
    _close(dz, g["dz_gq_f64"], 1e-9, 1e-13, "dz gq")
    _close(dmu, g["dmu_gq_f64"], 1e-9, 1e-13, "dmu gq")


@pytest.mark.parametrize("case", DEC_CASES)
def test_dec_f32_reference_within_1e6(golden, case):
    g = golden(case)
    _close(g["q"], g["q_f64"], 0, 1e-6, "q f32 vs truth")
    _close(g["p"], g["p_f64"], 0, 1e-6, "p f32 vs truth")


def test_target_distribution_sharded_colsum(golden):
    g = golden("dec_k4")
    q = g["q_f64"]
    f = q.sum(axis=0)
    halves = [dec_oracle.target_distribution(h, colsum=f) for h in (q[:30], q[30:])]
    _close(np.concatenate(halves), g["p_f64"], 1e-12, 0, "sharded p")


@pytest.mark.parametrize("dtag", ["f32", "f64"])
@pytest.mark.parametrize("k", [2, 4, 7])
def test_kmeans_lloyd_matches_sklearn(golden, dtag, k):
    g = golden("kmeans")
    X = g["X"].astype(np.float32 if dtag == "f32" else np.float64)
    res = kmeans_oracle.kmeans_fit(X, k, init=X[:k].copy(), n_init=1)
    pre = f"{dtag}_k{k}_"
    assert np.array_equal(res["labels"], g[pre + "labels"])
    assert res["n_iter"] == int(g[pre + "n_iter"])
    _close(res["centers"], g[pre + "centers"], 1e-5, 1e-5, "centres")
    _close(res["inertia"], g[pre + "inertia"], 1e-5, 0, "inertia")
    assert np.array_equal(kmeans_oracle.kmeans_predict(g["Xv"].astype(X.dtype), g[pre + "centers"]),
                          g[pre + "predict"])
    _close(kmeans_oracle.elbow_distortion(X, g[pre + "centers"]), g[pre + "elbow_train"], 1e-9, 0, "elbow")


def test_kmeans_empty_cluster_relocation(golden):
    g = golden("kmeans")
    res = kmeans_oracle.kmeans_fit(g["X"], 3, init=g["reloc_init"], n_init=1)
    assert np.array_equal(res["labels"], g["reloc_labels"])
    assert res["n_iter"] == int(g["reloc_n_iter"])
    _close(res["centers"], g["reloc_centers"], 1e-5, 1e-5, "centres")


def test_inertia_definitions(golden):
    g = golden("kmeans")
    a, X = g["f32_k4_labels"], g["X"]
    _close(kmeans_oracle.inertia_v1(a, X), g["inertia_v1_f32"], 1e-6, 0, "v1 f32")
    _close(kmeans_oracle.inertia_v2(a, X), g["inertia_v2_f32"], 1e-6, 0, "v2 f32")
    _close(kmeans_oracle.inertia_v1(a, X.astype(np.float64)), g["inertia_v1_f64"], 1e-10, 0, "v1 f64")
    _close(kmeans_oracle.inertia_v2(a, X.astype(np.float64)), g["inertia_v2_f64"], 1e-10, 0, "v2 f64")


@pytest.mark.parametrize("version", [1, 2])
def test_gap_statistic(golden, version):
    g = golden("gap")
    X = g["X"]

    def fit_predict(k, data):
        data = np.ascontiguousarray(data)
        return kmeans_oracle.kmeans_fit(data, k, init=data[:k].copy(), n_init=1)["labels"]

    np.random.seed(7)
    res = kmeans_oracle.gap_statistic(fit_predict, X, k_max=5, n_references=3, version=version)
    for col in ("k", "gap", "ref", "act", "ref_s"):
        _close(res[col], g[f"v{version}_{col}"], 1e-6, 1e-7, col)


@pytest.mark.parametrize("case", ["interp_c1", "interp_smoke", "interp_odd"])
def test_ref_port_matches_reference(golden, case):
    """The torch-CPU port timed as the CPU baseline reproduces the reference's float32 run."""
    import torch
    from oracle import ref_port
    g = golden(case)
    C = g["x"].shape[1] // 4
    R, H = int(g["R"]), float(g["hours"])
    sci, cci, rbf = ref_port.SingleChannelInterp(R, H, C), ref_port.CrossChannelInterp(C), \
        ref_port.RBFReadout(R, H, C)
    sci.kernel.data = torch.tensor(g["sci_kernel"])
    cci.kernel.data = torch.tensor(g["cci_kernel"])
    rbf.kernel.data = torch.tensor(g["rbf_kernel"])
    x = torch.tensor(g["x"])
    v = torch.tensor(g["v"], requires_grad=True)
    s = sci(x)
    c = cci(s)
    r = rbf(v, x)
    (c * torch.tensor(g["g_cci"])).sum().backward()
    (r * torch.tensor(g["g_rbf"])).sum().backward()
    _close(s.detach().numpy(), g["sci_out"], 1e-5, 1e-5, "sci")
    _close(c.detach().numpy(), g["cci_out"], 1e-5, 1e-5, "cci")
    _close(r.detach().numpy(), g["rbf_out"], 1e-5, 1e-5, "rbf")
    _close(sci.kernel.grad.numpy(), g["d_sci_kernel"], 1e-4, 1e-4, "d sci.kernel")
    _close(cci.kernel.grad.numpy(), g["d_cci_kernel"], 1e-4, 1e-4, "d cci.kernel")
    _close(rbf.kernel.grad.numpy(), g["d_rbf_kernel"], 1e-4, 1e-4, "d rbf.kernel")
    _close(v.grad.numpy(), g["dv"], 1e-4, 1e-5, "dv")


def test_ref_port_dec(golden):
    import torch
    from oracle import ref_port
    g = golden("dec_k4")
    z = torch.tensor(g["z"], requires_grad=True)
    mu = torch.tensor(g["mu"], requires_grad=True)
    loss, q, p = ref_port.dec_step(z, mu, float(g["alpha"]))
    _close(q.numpy(), g["q"], 0, 1e-6, "q")
    _close(p.numpy(), g["p"], 0, 1e-6, "p")
    _close(loss.numpy(), g["kl"], 1e-5, 1e-8, "kl")
    _close(z.grad.numpy(), g["dz_kl"], 1e-4, 1e-8, "dz")
    _close(mu.grad.numpy(), g["dmu_kl"], 1e-4, 1e-7, "dmu")


@pytest.mark.parametrize("tag", ["blobs", "dups"])
def test_internal_metrics_oracle_matches_reference(golden, tag):
    """Dunn / Calinski-Harabasz / Davies-Bouldin restatements against the values the reference's own classes gave
    (internal_eval.py:15-147; fixture internal_eval.npz from oracle/gen_golden.py, incl. the duplicated-row case
    where a touching pair of clusters drops out of the Dunn minimum)."""
    g = golden("internal_eval")
    X, lab = g[tag + "_X"], g[tag + "_labels"]
    np.testing.assert_allclose(kmeans_oracle.dunn_index(X, lab), float(g[tag + "_dunn"]), rtol=1e-6)
    # sklearn evaluates float32 input partly in float32 (centroid distances): float32-rounding agreement
    np.testing.assert_allclose(kmeans_oracle.calinski_harabasz(X, lab), float(g[tag + "_ch"]), rtol=1e-6)
    np.testing.assert_allclose(kmeans_oracle.davies_bouldin(X, lab), float(g[tag + "_db"]), rtol=1e-6)
