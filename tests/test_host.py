"""CPU-side checks: the C-ABI library loads and exports every declared symbol, the nn.Module
mirrors keep the reference's interface and state-dict keys, there is no CPU fallback, and the
drop-in shim makes the reference's own model files resolve to the B200 operators."""
import os
import re
import subprocess
import sys
import types

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"


def _declared_symbols():
    text = open(os.path.join(REPO, "include", "dic_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dic_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from deep_interpolation_clustering_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "run `python -m deep_interpolation_clustering_b200.build` first"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True,
                         check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if line.strip()}
    declared = _declared_symbols()
    assert len(declared) >= 20
    assert [s for s in declared if s not in exported] == []
    assert sorted(_lib.SIGNATURES) == declared          # the ctypes table mirrors the header 1:1
    lib = _lib.lib()                                     # loads without a GPU
    assert lib.dic_version() == 100


def test_library_built_for_sm100a_with_tma():
    """The shipped cubin is sm_100a and the staging path really is a TMA bulk copy (UBLKCP)."""
    from deep_interpolation_clustering_b200 import _lib
    r = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in r.stdout
    obj = os.path.join(os.path.dirname(_lib.LIB_PATH), "csrc", "build", "interp_sci.o")
    if not os.path.exists(obj):
        pytest.skip("object cache not present")
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass and "MUFU.EX2" in sass and "SYNCS.ARRIVE.TRANS64" in sass


def test_argument_errors_without_gpu():
    """Validation happens before any CUDA call, so it is testable on a CPU box."""
    from deep_interpolation_clustering_b200 import _lib
    lib = _lib.lib()
    assert lib.dic_sci_fwd(None, None, None, None, None, 1, 6, 64, 48, 0, None) == _lib.DIC_ERR_INVALID_ARGUMENT
    assert "null" in _lib.last_error()
    # T so large that one encounter cannot be staged in 227 KB of shared memory
    assert lib.dic_sci_fwd(16, 16, 16, 16, None, 1, 6, 100000, 48, 0, None) == _lib.DIC_ERR_UNSUPPORTED
    assert "shared memory" in _lib.last_error()
    assert lib.dic_cci_fwd(16, 16, 16, 1, 17, 48, None) == _lib.DIC_ERR_UNSUPPORTED
    assert lib.dic_dec_q_fwd(16, 16, 16, None, None, None, 4, 30, 4, 1.0, None) == _lib.DIC_ERR_UNSUPPORTED
    assert lib.dic_dec_q_fwd(16, 16, 16, None, None, None, 4, 32, 4, -1.0, None) == _lib.DIC_ERR_INVALID_ARGUMENT
    assert lib.dic_kmeans_assign(16, 16, 16, None, None, 16, 16, 10, 8, 100, 0, 0, None) == _lib.DIC_ERR_UNSUPPORTED
    with pytest.raises(ValueError):
        _lib.check(_lib.DIC_ERR_UNSUPPORTED, "x")
    with pytest.raises(_lib.DicError):
        _lib.check(_lib.DIC_ERR_CUDA, "x")


def test_module_interfaces_and_state_dict_keys():
    import deep_interpolation_clustering_b200 as dic
    dev = torch.device("cpu")
    sci = dic.SingleChannelInterp(48, 24, 6, 64, dev)
    cci = dic.CrossChannelInterp(6, 64, dev)
    rbf = dic.RBF(24, 48, 256, 6, 0.2, dic.basis_func_dict()["gaussian"], dev)
    ca = dic.ClusterAssignment(4, 256, 1.0)
    assert tuple(sci.kernel.shape) == (6,) and 0 <= float(sci.kernel.min()) and float(sci.kernel.max()) < 1
    assert torch.equal(cci.kernel.data, torch.eye(6))
    assert list(sci.state_dict()) == ["kernel"] and list(cci.state_dict()) == ["kernel"]
    assert set(rbf.state_dict()) == {
        "kernel", "compress_fc.module.model.0.weight", "compress_fc.module.model.0.bias",
        "compress_fc.module.model.1.weight", "compress_fc.module.model.1.bias",
        "compress_fc.module.model.1.running_mean", "compress_fc.module.model.1.running_var",
        "compress_fc.module.model.1.num_batches_tracked", "compress_fc.module.model.4.weight",
        "compress_fc.module.model.4.bias"}
    assert "interp_t" not in rbf.state_dict() and tuple(rbf.interp_t.shape) == (48,)
    assert list(ca.state_dict()) == ["cluster_centers"] and tuple(ca.cluster_centers.shape) == (4, 256)
    c = torch.randn(4, 256)
    ca.init_center(c)
    assert ca.get_center() is ca.cluster_centers and torch.equal(ca.cluster_centers.data, c)
    with pytest.raises(NotImplementedError):
        dic.RBF(24, 48, 256, 6, 0.2, lambda b, a: a, dev)


def test_no_cpu_fallback():
    import deep_interpolation_clustering_b200 as dic
    dev = torch.device("cpu")
    sci = dic.SingleChannelInterp(8, 24, 6, 16, dev)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sci(torch.zeros(2, 24, 16))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dic.CrossChannelInterp(6, 16, dev)(torch.zeros(2, 8, 18))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dic.ClusterAssignment(4, 64)(torch.zeros(8, 64))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dic.target_distribution(torch.full((8, 4), 0.25))
    if not torch.cuda.is_available():
        from deep_interpolation_clustering_b200.kmeans import KMeansB200
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            KMeansB200(n_clusters=2).fit(np.zeros((10, 4), np.float32))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(REPO, "deep_interpolation_clustering_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "oracle" not in src.replace("oracle/gen_golden.py", ""), f"{f} mentions the oracle"


def test_same_clustering_and_random_state_helpers():
    from deep_interpolation_clustering_b200.kmeans import _check_random_state, _same_clustering
    assert _same_clustering([0, 0, 1, 2], [2, 2, 0, 1], 3)
    assert not _same_clustering([0, 0, 1, 2], [2, 1, 0, 1], 3)
    assert _check_random_state(None) is np.random.mtrand._rand
    assert _check_random_state(3).uniform() == np.random.RandomState(3).uniform()


def test_restart_comparison_on_tensors_equals_the_host_rule():
    """KMeansB200._same_as_best (K x K contingency table where the labels live) against _same_clustering
    (sklearn's _is_same_clustering restated) on random pairs of label vectors: equal, permuted, merged, split."""
    import torch
    from deep_interpolation_clustering_b200.kmeans import KMeansB200, _Comm, _same_clustering
    comm = _Comm.__new__(_Comm)
    comm.group, comm.on, comm.rank, comm.size = None, False, 0, 1
    km = KMeansB200(n_clusters=4)
    rng = np.random.RandomState(5)
    seen = set()
    for trial in range(60):
        K = int(rng.randint(2, 7))
        a = rng.randint(0, K, size=200)
        mode = trial % 4
        if mode == 0:
            b = a.copy()
        elif mode == 1:
            b = rng.permutation(K)[a]
        elif mode == 2:
            b = np.where(a == K - 1, 0, a)                      # two clusters merged
        else:
            b = np.where((a == 0) & (rng.uniform(size=a.size) < 0.5), K - 1, a)   # one cluster split
        want = _same_clustering(a, b, K)
        got = km._same_as_best(torch.from_numpy(a.astype(np.int32)), torch.from_numpy(b.astype(np.int32)), K, comm)
        assert got == want, (trial, mode, K)
        seen.add(want)
    assert seen == {True, False}


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout not present")
def test_dropin_makes_reference_models_use_b200_operators():
    """pretrain_interp.Net / clustering_interp.Net (unchanged upstream files) build on top of the
    B200 mirrors and expose exactly the reference's state-dict keys."""
    from deep_interpolation_clustering_b200 import dropin
    import deep_interpolation_clustering_b200 as dic

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        return m

    saved = {k: sys.modules.get(k) for k in ("tensorflow", "warmup_scheduler", "utils", "pretrain_interp",
                                             "clustering_interp", "interpolation_layer", "rbf", "dec", "info")}
    sys.modules.setdefault("tensorflow", stub("tensorflow", random=types.SimpleNamespace(set_seed=lambda s: None)))
    sys.modules.setdefault("warmup_scheduler", stub("warmup_scheduler", GradualWarmupScheduler=object))
    sys.path.insert(0, REFERENCE)
    cwd = os.getcwd()
    os.chdir("/tmp")
    try:
        args = types.SimpleNamespace(num_variables=6, num_timestamps=64, ref_points=48, hours_from_admission=24,
                                     dropout=0.2, aux_tasks={}, fake_detection=False, triple_margin=0.,
                                     cluster_number=4)
        for k in ("pretrain_interp", "clustering_interp", "interpolation_layer", "rbf", "dec"):
            sys.modules.pop(k, None)
        import clustering_interp as ref_ci                      # the reference's own operators
        ref_keys = list(ref_ci.Net(args, torch.device("cpu")).state_dict())
        for k in ("pretrain_interp", "clustering_interp", "interpolation_layer", "rbf", "dec"):
            sys.modules.pop(k, None)
        dropin.install()
        import clustering_interp as new_ci
        import pretrain_interp as new_pi
        net = new_ci.Net(args, torch.device("cpu"))
        assert isinstance(net.sci, dic.SingleChannelInterp) and isinstance(net.cci, dic.CrossChannelInterp)
        assert isinstance(net.rbf, dic.RBF) and isinstance(net.cluster_assignment, dic.ClusterAssignment)
        assert list(net.state_dict()) == ref_keys
        assert isinstance(new_pi.Net(args, torch.device("cpu")).sci, dic.SingleChannelInterp)
    finally:
        dropin.uninstall()
        os.chdir(cwd)
        sys.path.remove(REFERENCE)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout not present")
def test_patch_trainer_routes_the_real_trainer_class_through_epoch_path():
    """dropin.patch_trainer on the reference's own clustering_trainer.TrainerCluster (unchanged upstream file):
    generate_pred_cluster (clustering_trainer.py:473-484) and merge_ob_pred (:486-493) of the patched class return what
    the reference's methods return on the same batches, with the epoch's latents accumulated by the forward hook."""
    from deep_interpolation_clustering_b200 import dropin
    from deep_interpolation_clustering_b200.kmeans import KMeansB200

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        return m

    names = ("tensorflow", "warmup_scheduler", "tensorboardX", "utils", "clustering_trainer", "clustering_interp",
             "interpolation_layer", "rbf", "dec", "info", "dataloader")
    saved = {k: sys.modules.get(k) for k in names}
    sys.modules.setdefault("tensorflow", stub("tensorflow", random=types.SimpleNamespace(set_seed=lambda s: None)))
    sys.modules.setdefault("warmup_scheduler", stub("warmup_scheduler", GradualWarmupScheduler=object))
    sys.modules.setdefault("tensorboardX", stub("tensorboardX", SummaryWriter=object))
    sys.path.insert(0, REFERENCE)
    cwd = os.getcwd()
    os.chdir("/tmp")
    try:
        for k in ("clustering_trainer", "clustering_interp", "interpolation_layer", "rbf", "dec"):
            sys.modules.pop(k, None)
        import clustering_trainer as ct
        ref_generate = ct.TrainerCluster.generate_pred_cluster
        ref_merge = ct.TrainerCluster.merge_ob_pred

        class Model(torch.nn.Module):                       # (hidden, rec_ob, aux_pred_dict) like clustering_interp.Net
            def forward(self, x):
                q = torch.softmax(x[:, :4] * 3.0, dim=1)
                return x * 2.0, x, {"cluster_pred": q, "cluster_label": q.detach()}

        rng = np.random.RandomState(0)
        batches = [torch.from_numpy(rng.normal(size=(n, 8)).astype(np.float32)) for n in (5, 7, 3)]

        def fake_eval(self, scope, dl, denoise=False):      # the data path of eval_one_epoch (:285-422), nothing else
            lst = []
            for b in dl:
                hidden, rec, aux = self.model(b)
                d = {"encounter_id": [f"e{i}" for i in range(b.shape[0])], "ob": b.numpy(),
                     "hidden": hidden.detach().cpu().numpy(), "rec_ob": rec.detach().cpu().numpy()}
                d.update({k: v.detach().cpu().numpy() for k, v in aux.items()})
                lst.append(d)
            return {"scope": scope}, lst

        ct.TrainerCluster.eval_one_epoch = fake_eval
        ref_tr = object.__new__(ct.TrainerCluster)
        ref_tr.model = Model()
        prev = np.array([0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2])
        want_first = ref_generate(ref_tr, "valid", batches, None)
        want = ref_generate(ref_tr, "valid", batches, prev)
        want_merged = ref_merge(ref_tr, fake_eval(ref_tr, "valid", batches)[1])

        cls = dropin.patch_trainer(ct)
        assert ct.KMeans is KMeansB200 and cls is ct.TrainerCluster
        tr = object.__new__(cls)
        tr.model = Model()
        got_first = tr.generate_pred_cluster("valid", batches, None)
        got = tr.generate_pred_cluster("valid", batches, prev)
        assert got_first[0] == want_first[0] == 1.0
        assert got[0] == want[0] and np.array_equal(got[1], want[1]) and got[2] == want[2]
        feats = tr.dic_epoch_features
        assert feats.n == 15 and torch.equal(feats.features(), torch.cat(batches) * 2.0)
        merged = tr.merge_ob_pred(tr.eval_one_epoch("valid", batches)[1])
        assert sorted(merged) == sorted(want_merged)
        for k, v in want_merged.items():
            assert np.array_equal(merged[k], v), k
            assert merged[k].dtype == v.dtype or v.dtype.kind in "US", k
    finally:
        os.chdir(cwd)
        sys.path.remove(REFERENCE)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
