#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE ITSELF on seeded inputs.

TEST INFRASTRUCTURE.  Run in the build container only (needs /root/reference):

    python oracle/gen_golden.py [--out tests/golden]

The reference modules are imported from /root/reference through stub modules for
its absent optional imports (tensorflow, warmup_scheduler, seaborn, matplotlib,
kneed) - SURVEY.md Appendix C.  Nothing of the reference is copied: the fixtures
hold only inputs and the numbers the reference produced (float32 as shipped, and
the same modules run in float64 as the rounding-free "truth").

Versions are recorded in tests/golden/MANIFEST.json.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import types
import warnings

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
REFERENCE = os.environ.get("DIC_REFERENCE", "/root/reference")


def import_reference():
    """Put the reference on sys.path with stubs for imports the container lacks."""
    def stub(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m
    for name, attrs in (
        ("tensorflow", dict(random=types.SimpleNamespace(set_seed=lambda s: None))),
        ("warmup_scheduler", dict(GradualWarmupScheduler=object)),
        ("seaborn", {}), ("kneed", dict(KneeLocator=object)),
    ):
        try:
            __import__(name)
        except Exception:
            stub(name, **attrs)
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        mp = stub("matplotlib")
        mp.pyplot = stub("matplotlib.pyplot")
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    argv, sys.argv = sys.argv, ["x"]
    cwd = os.getcwd()
    os.chdir("/tmp")          # the reference's logger writes a file relative to cwd
    try:
        import interpolation_layer, dec, rbf, p2_clustering_optK, internal_eval  # noqa: E401
    finally:
        sys.argv = argv
        os.chdir(cwd)
    return types.SimpleNamespace(interpolation_layer=interpolation_layer, dec=dec, rbf=rbf,
                                 p2=p2_clustering_optK, internal_eval=internal_eval)


def _interp_case(ref, x, params, hours, R, seed, backward=True):
    import torch
    B, C4, T = x.shape
    C = C4 // 4
    rng = np.random.RandomState(seed)
    v = rng.normal(size=(B, C, R)).astype(np.float32)
    g_cci = rng.normal(size=(B, R, 3 * C)).astype(np.float32)
    g_rbf = rng.normal(size=(B, C, T)).astype(np.float32)
    out = dict(x=x, v=v, g_cci=g_cci, g_rbf=g_rbf, hours=np.float64(hours), R=np.int64(R),
               **{k: np.asarray(a) for k, a in params.items()})
    dev = torch.device("cpu")
    for tag, dt in (("", torch.float32), ("_f64", torch.float64)):
        sci = ref.interpolation_layer.SingleChannelInterp(R, hours, C, T, dev)
        cci = ref.interpolation_layer.CrossChannelInterp(C, T, dev)
        rbf = ref.rbf.RBF(hours, R, C, C, 0.0, ref.rbf.basis_func_dict()["gaussian"], dev)
        rbf.compress_fc = torch.nn.Identity()          # boundary of the custom kernel is v
        sci.kernel.data = torch.tensor(params["sci_kernel"])
        cci.kernel.data = torch.tensor(params["cci_kernel"])
        rbf.kernel.data = torch.tensor(params["rbf_kernel"])
        if dt == torch.float64:
            sci, cci, rbf = sci.double(), cci.double(), rbf.double()
            rbf.interp_t = rbf.interp_t.double()
            # float64 grid values = the float32 linspace values, so both runs share inputs
            _lin = torch.linspace
            torch.linspace = lambda *a, **k: _lin(*a, **k).double()
        else:
            out["ref_t"] = torch.linspace(0, hours, R).numpy()
        try:
            xt = torch.tensor(x, dtype=dt)
            vt = torch.tensor(v, dtype=dt, requires_grad=True)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                s = sci(xt)
                s.retain_grad()
                c = cci(s)
                r = rbf(vt, xt)
                out["sci_out" + tag] = s.detach().contiguous().numpy()
                out["cci_out" + tag] = c.detach().contiguous().numpy()
                out["rbf_out" + tag] = r.detach().numpy()
                if backward:
                    (c * torch.tensor(g_cci, dtype=dt)).sum().backward()
                    (r * torch.tensor(g_rbf, dtype=dt)).sum().backward()
                    out["d_sci_kernel" + tag] = sci.kernel.grad.numpy()
                    out["d_cci_kernel" + tag] = cci.kernel.grad.numpy()
                    out["d_rbf_kernel" + tag] = rbf.kernel.grad.numpy()
                    out["d_sci_out" + tag] = s.grad.contiguous().numpy()
                    out["dv" + tag] = vt.grad.numpy()
        finally:
            if dt == torch.float64:
                torch.linspace = _lin
    return out


def gen_interp(ref, outdir, only=None):
    from deep_interpolation_clustering_b200 import synth
    # trained kernels leave U[0,1): very wide (softplus(-6) = 0.0025: the window is the whole record) and very narrow
    # (softplus(8) = 8: a weight falls below 2^-27 within 1.5 h) Gaussians exercise the window cut-offs of the
    # CUDA kernels (kCut, kRbfCut) where they bite
    extreme = dict(sci_kernel=np.array([-6.0, -2.0, 0.5, 3.0, 8.0, 1.0], np.float32),
                   rbf_kernel=np.array([8.0, 3.0, -2.0, -6.0, 0.3, 5.0], np.float32))
    cases = {
        "interp_kernels": (synth.make_encounters(6, 6, 128, 24.0, seed=80), 24.0, 96, extreme),
        "interp_c1": (synth.make_encounters(8, 6, 64, 24.0, seed=0), 24.0, 48),
        "interp_c2": (synth.make_encounters(4, 6, 256, 24.0, seed=10), 24.0, 96),
        "interp_c5": (synth.make_encounters(2, 6, 1024, 24.0, seed=20), 24.0, 192),
        "interp_smoke": (synth.make_adversarial_encounters(10, 6, 30, 6.0, seed=30), 6.0, 11),
        "interp_odd": (synth.make_adversarial_encounters(5, 3, 21, 24.0, seed=40), 24.0, 7),
        "interp_dense": (synth.make_encounters(3, 6, 64, 24.0, seed=50, min_obs=64), 24.0, 48),
    }
    for name, case in cases.items():
        if only and name not in only:
            continue
        x, hours, R = case[:3]
        C = x.shape[1] // 4
        params = synth.make_interp_params(C, seed=1)
        if len(case) > 3:
            params.update(case[3])
        np.savez_compressed(os.path.join(outdir, name + ".npz"),
                            **_interp_case(ref, x, params, hours, R, seed=2))
    if only and "interp_allmasked" not in only:
        return
    # all-masked channel: reference yields w=-inf, y=NaN (forward only)
    x = synth.make_encounters(3, 6, 16, 24.0, seed=60)
    x[1, 0:6][2] = 0; x[1, 6:12][2] = 0; x[1, 12:18][2] = 0
    np.savez_compressed(os.path.join(outdir, "interp_allmasked.npz"),
                        **_interp_case(ref, x, synth.make_interp_params(6, seed=1), 24.0, 12,
                                       seed=2, backward=False))


def gen_rbf_module(ref, outdir):
    """Full RBF module (compress_fc included, eval mode) for the state-dict drop-in test."""
    import torch
    from deep_interpolation_clustering_b200 import synth
    torch.manual_seed(5)
    B, C, T, R, H, IN = 6, 6, 64, 48, 24.0, 256
    x = synth.make_encounters(B, C, T, H, seed=70)
    rbf = ref.rbf.RBF(H, R, IN, C, 0.2, ref.rbf.basis_func_dict()["gaussian"], torch.device("cpu"))
    bn = rbf.compress_fc.module.model[1]
    bn.running_mean.data = 0.1 * torch.randn(128)
    bn.running_var.data = 1.0 + 0.2 * torch.rand(128)
    rbf.eval()
    interp = torch.randn(B, IN, R)
    with torch.no_grad():
        out = rbf(interp, torch.tensor(x))
    sd = {"sd." + k: v.numpy() for k, v in rbf.state_dict().items()}
    np.savez_compressed(os.path.join(outdir, "rbf_module.npz"), x=x, interp=interp.numpy(),
                        out=out.numpy(), hours=np.float64(H), R=np.int64(R), **sd)


def gen_dec(ref, outdir):
    import torch
    import torch.nn.functional as F
    from deep_interpolation_clustering_b200 import synth
    for name, (B, D, K, alpha) in {"dec_k4": (64, 256, 4, 1.0), "dec_k16": (50, 64, 16, 1.0),
                                   "dec_alpha2": (33, 40, 3, 2.0), "dec_k10": (40, 64, 10, 1.0)}.items():
        z, mu = synth.make_latents(B, D, K, seed=3)
        mu = mu + 0.05 * np.random.RandomState(4).normal(size=mu.shape).astype(np.float32)
        gq = np.random.RandomState(5).normal(size=(B, K)).astype(np.float32)
        out = dict(z=z, mu=mu, gq=gq, alpha=np.float64(alpha))
        for tag, dt in (("", torch.float32), ("_f64", torch.float64)):
            ca = ref.dec.ClusterAssignment(K, D, alpha, torch.tensor(mu, dtype=dt))
            zt = torch.tensor(z, dtype=dt, requires_grad=True)
            q = ca(zt)
            p = ref.dec.target_distribution(q).detach()
            kl = F.kl_div(q.log(), p, reduction="batchmean")      # clustering_interp.py:205-207
            kl.backward()
            out.update({"q" + tag: q.detach().numpy(), "p" + tag: p.numpy(),
                        "kl" + tag: kl.detach().numpy(), "dz_kl" + tag: zt.grad.numpy().copy(),
                        "dmu_kl" + tag: ca.cluster_centers.grad.numpy().copy()})
            zt.grad = None
            ca.cluster_centers.grad = None
            (ca(zt) * torch.tensor(gq, dtype=dt)).sum().backward()
            out.update({"dz_gq" + tag: zt.grad.numpy().copy(),
                        "dmu_gq" + tag: ca.cluster_centers.grad.numpy().copy()})
        np.savez_compressed(os.path.join(outdir, name + ".npz"), **out)


class FixedInitKMeans:
    """Duck type handed to the reference's gap routine: deterministic init = first k rows,
    so no RNG stream has to be matched."""

    def __init__(self):
        self.n_clusters = 2

    def fit_predict(self, X):
        from sklearn.cluster import KMeans
        X = np.ascontiguousarray(X)
        return KMeans(n_clusters=self.n_clusters, init=X[:self.n_clusters].copy(), n_init=1).fit_predict(X)


def gen_kmeans(ref, outdir):
    from sklearn.cluster import KMeans
    from scipy.spatial.distance import cdist
    from deep_interpolation_clustering_b200 import synth
    out = {}
    X32 = synth.make_blobs(1500, 16, 5, seed=0)
    Xv = synth.make_blobs(400, 16, 5, seed=0)[::-1].copy()
    out["X"], out["Xv"] = X32, Xv
    for dtag, X in (("f32", X32), ("f64", X32.astype(np.float64))):
        for k in (2, 4, 7):
            init = X[:k].copy()
            km = KMeans(n_clusters=k, init=init, n_init=1).fit(X)
            pre = f"{dtag}_k{k}_"
            out[pre + "labels"] = km.labels_
            out[pre + "centers"] = km.cluster_centers_
            out[pre + "inertia"] = np.float64(km.inertia_)
            out[pre + "n_iter"] = np.int64(km.n_iter_)
            out[pre + "predict"] = km.predict(Xv.astype(X.dtype))
            out[pre + "elbow_train"] = np.float64(
                sum(np.min(cdist(X, km.cluster_centers_, "euclidean"), axis=1)) / X.shape[0])
    # empty-cluster relocation: one initial centre far away from all data
    init = np.concatenate([X32[:2], np.full((1, 16), 1e3, np.float32)])
    km = KMeans(n_clusters=3, init=init, n_init=1).fit(X32)
    out["reloc_init"] = init
    out["reloc_labels"], out["reloc_centers"] = km.labels_, km.cluster_centers_
    out["reloc_inertia"], out["reloc_n_iter"] = np.float64(km.inertia_), np.int64(km.n_iter_)
    # the reference's own inertia definitions
    kmobj = ref.p2.KM(5, "/tmp/dic_golden_km", ["Sihouette", "Davies-Bouldin_Index", "Calinski-Harabasz"], 1, 3)
    a = out["f32_k4_labels"]
    out["inertia_v1_f32"] = np.float64(kmobj.compute_inertia_v1(a, X32))
    out["inertia_v2_f32"] = np.float64(kmobj.computer_intertia_v2(a, X32))
    out["inertia_v1_f64"] = np.float64(kmobj.compute_inertia_v1(a, X32.astype(np.float64)))
    out["inertia_v2_f64"] = np.float64(kmobj.computer_intertia_v2(a, X32.astype(np.float64)))
    np.savez_compressed(os.path.join(outdir, "kmeans.npz"), **out)

    # gap statistic + internal metrics through the reference's own routine
    Xg = synth.make_blobs(300, 8, 3, seed=11)
    res = {"X": Xg}
    for version in (1, 2):
        np.random.seed(7)
        df = kmobj.compute_gap_internal_metric(FixedInitKMeans(), Xg, k_max=5, n_references=3,
                                               version=version)
        for col in df.columns:
            res[f"v{version}_{col}"] = df[col].to_numpy(dtype=np.float64)
    np.savez_compressed(os.path.join(outdir, "gap.npz"), **res)


def gen_internal_eval(ref, outdir):
    """The four cluster-validity metrics of internal_eval.py:15-147 through the reference's own classes (Dunn's
    pure-Python double loop limits the size): blobs, a k-means labelling, and a second data set with duplicated rows
    (zero inter-cluster distances, which internal_eval.py:104 skips through ``nonzero()``)."""
    from sklearn.cluster import KMeans
    from deep_interpolation_clustering_b200 import synth
    out = {}
    X = synth.make_blobs(700, 16, 4, seed=21)
    lab = KMeans(n_clusters=5, init=X[:5].copy(), n_init=1).fit_predict(X)
    Xd = synth.make_blobs(240, 8, 3, seed=22)
    labd = KMeans(n_clusters=3, init=Xd[:3].copy(), n_init=1).fit_predict(Xd)
    src = np.nonzero(labd[:200] == 0)[0][:12]           # 12 rows of cluster 0 ...
    Xd[200:212] = Xd[src]                               # ... duplicated exactly ...
    labd[200:206] = 1                                   # ... 6 copies filed under cluster 1 (the pair (0,1) then has a
    labd[206:212] = 0                                   # zero nearest distance and drops out), 6 under cluster 0
    for tag, (A, a) in (("blobs", (X, lab)), ("dups", (Xd, labd))):
        out[tag + "_X"], out[tag + "_labels"] = A, a.astype(np.int64)
        out[tag + "_dunn"] = np.float64(ref.internal_eval.DunnIndex()(A, a))
        out[tag + "_silhouette"] = np.float64(ref.internal_eval.Sihouette()(A, a))
        out[tag + "_ch"] = np.float64(ref.internal_eval.CHIndex()(A, a))
        out[tag + "_db"] = np.float64(ref.internal_eval.DBIndex()(A, a))
    np.savez_compressed(os.path.join(outdir, "internal_eval.npz"), **out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(REPO, "tests", "golden"))
    ap.add_argument("--only", nargs="*", help="regenerate only these interpolation fixtures (leaves the others untouched)")
    ap.add_argument("--internal-eval-only", action="store_true", help="regenerate only internal_eval.npz")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    ref = import_reference()
    if not args.internal_eval_only:
        gen_interp(ref, args.out, args.only)
    if not args.only and not args.internal_eval_only:
        gen_rbf_module(ref, args.out)
        gen_dec(ref, args.out)
        gen_kmeans(ref, args.out)
    if not args.only:
        gen_internal_eval(ref, args.out)
    import torch, sklearn, scipy
    manifest = dict(reference=REFERENCE, torch=torch.__version__, numpy=np.__version__,
                    sklearn=sklearn.__version__, scipy=scipy.__version__,
                    generator="oracle/gen_golden.py",
                    files=sorted(f for f in os.listdir(args.out) if f.endswith(".npz")))
    with open(os.path.join(args.out, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    print(json.dumps(manifest, indent=1))


if __name__ == "__main__":
    main()
