"""GPU parity of the interpolation network (SCI / CCI / RBF) against the oracle.

Every call goes nn.Module mirror -> autograd Function -> ctypes -> C ABI -> sm_100a kernel.
Truth = the reference run in float64 (tests/golden, written by oracle/gen_golden.py) or the
numpy oracle in float64 on seeded inputs; tolerance = rtol 1e-5 (BASELINE.json) plus an
absolute floor of 1e-5 x the tensor's RMS for entries that cancel to ~0.
"""
import numpy as np
import pytest
import torch

from gpu_util import RTOL_GRID, record, scale_atol

pytestmark = pytest.mark.gpu

CASES = ["interp_c1", "interp_c2", "interp_c5", "interp_smoke", "interp_odd", "interp_dense", "interp_kernels"]


def _modules(g, dev):
    import deep_interpolation_clustering_b200 as dic
    x = g["x"]
    C, T = x.shape[1] // 4, x.shape[2]
    R, H = int(g["R"]), float(g["hours"])
    sci = dic.SingleChannelInterp(R, H, C, T, dev)
    cci = dic.CrossChannelInterp(C, T, dev)
    rbf = dic.RBF(H, R, C, C, 0.0, dic.basis_func_dict()["gaussian"], dev)
    rbf.compress_fc = torch.nn.Identity()       # golden v is the kernel boundary
    sci.kernel.data = torch.tensor(g["sci_kernel"], device=dev)
    cci.kernel.data = torch.tensor(g["cci_kernel"], device=dev)
    rbf.kernel.data = torch.tensor(g["rbf_kernel"], device=dev)
    return sci, cci, rbf


def _check(name, got, truth, rtol=RTOL_GRID):
    return record(name, got.detach().cpu().numpy(), truth, rtol, scale_atol(truth, rtol))


def _check_groups(name, got, truth, C, rtol=RTOL_GRID):
    """(B, R, 3C) outputs: the three channel groups have very different magnitudes (the
    log-intensity w reaches hundreds), so each group gets its own RMS-scaled floor."""
    got = got.detach().cpu().numpy()
    for i, grp in enumerate(("a", "b", "c")):
        sl = slice(i * C, (i + 1) * C)
        record(f"{name}[{grp}]", got[..., sl], truth[..., sl], rtol, scale_atol(truth[..., sl], rtol))


@pytest.mark.parametrize("case", CASES)
def test_forward_matches_reference_f64(golden, case):
    g = golden(case)
    dev = torch.device("cuda:0")
    sci, cci, rbf = _modules(g, dev)
    x = torch.tensor(g["x"], device=dev)
    s = sci(x)
    assert s.shape == g["sci_out"].shape
    C, R = g["x"].shape[1] // 4, int(g["R"])
    assert s.stride() == (3 * C * R, 1, R)            # the reference's permuted view
    c = cci(s)
    r = rbf(torch.tensor(g["v"], device=dev), x)
    _check_groups(f"{case}/sci", s, g["sci_out_f64"], C)
    _check_groups(f"{case}/cci", c, g["cci_out_f64"], C)
    _check(f"{case}/rbf", r, g["rbf_out_f64"])
    # and within the float32 reference's own error of the shipped float32 numbers
    _check(f"{case}/sci_vs_f32ref", s, g["sci_out"], 1e-4)
    _check(f"{case}/cci_vs_f32ref", c, g["cci_out"], 1e-4)
    _check(f"{case}/rbf_vs_f32ref", r, g["rbf_out"], 1e-4)


@pytest.mark.parametrize("case", CASES)
def test_backward_matches_reference_f64(golden, case):
    g = golden(case)
    dev = torch.device("cuda:0")
    sci, cci, rbf = _modules(g, dev)
    x = torch.tensor(g["x"], device=dev)
    v = torch.tensor(g["v"], device=dev, requires_grad=True)
    s = sci(x)
    s.retain_grad()
    c = cci(s)
    r = rbf(v, x)
    (c * torch.tensor(g["g_cci"], device=dev)).sum().backward()
    (r * torch.tensor(g["g_rbf"], device=dev)).sum().backward()
    _check_groups(f"{case}/d_sci_out", s.grad, g["d_sci_out_f64"], g["x"].shape[1] // 4)
    _check(f"{case}/d_cci_kernel", cci.kernel.grad, g["d_cci_kernel_f64"])
    _check(f"{case}/d_sci_kernel", sci.kernel.grad, g["d_sci_kernel_f64"])
    _check(f"{case}/dv", v.grad, g["dv_f64"])
    _check(f"{case}/d_rbf_kernel", rbf.kernel.grad, g["d_rbf_kernel_f64"])


def test_allmasked_channel_semantics(golden):
    """w = -inf, y = y' = NaN for an all-masked channel, like the reference."""
    g = golden("interp_allmasked")
    dev = torch.device("cuda:0")
    sci, cci, _ = _modules(g, dev)
    x = torch.tensor(g["x"], device=dev)
    s = sci(x)
    _check("allmasked/sci", s, g["sci_out_f64"])
    _check("allmasked/cci", cci(s), g["cci_out_f64"])


def test_config1_full_size_vs_oracle():
    """BASELINE config 1: 1,000 encounters x 6 vitals x <=64 obs, 48 reference points."""
    from deep_interpolation_clustering_b200 import synth
    import deep_interpolation_clustering_b200 as dic
    from oracle import interp_oracle
    dev = torch.device("cuda:0")
    B, C, T, R, H = 1000, 6, 64, 48, 24.0
    xn = synth.make_encounters(B, C, T, H, seed=0)
    p = synth.make_interp_params(C, seed=1)
    rng = np.random.RandomState(2)
    vn = rng.normal(size=(B, C, R)).astype(np.float32)
    gc = rng.normal(size=(B, R, 3 * C)).astype(np.float32)
    rt = interp_oracle.linspace_grid(H, R)
    x64 = xn.astype(np.float64)
    s64 = interp_oracle.sci_forward(x64, p["sci_kernel"].astype(np.float64), rt, C)
    c64 = interp_oracle.cci_forward(s64, p["cci_kernel"].astype(np.float64), C)
    r64 = interp_oracle.rbf_forward(vn, x64, p["rbf_kernel"].astype(np.float64), rt, C)
    du64, dK64 = interp_oracle.cci_backward(s64, p["cci_kernel"].astype(np.float64), C, gc)
    dk64 = interp_oracle.sci_backward(x64, p["sci_kernel"].astype(np.float64), rt, C, du64)
    grec = interp_oracle.rec_loss_grad(x64, r64, C)
    dv64, dkr64 = interp_oracle.rbf_backward(vn, x64, p["rbf_kernel"].astype(np.float64), rt, C, grec)
    loss64 = interp_oracle.rec_loss(x64, r64, C)

    sci = dic.SingleChannelInterp(R, H, C, T, dev)
    cci = dic.CrossChannelInterp(C, T, dev)
    rbf = dic.RBF(H, R, C, C, 0.0, dic.basis_func_dict()["gaussian"], dev)
    rbf.compress_fc = torch.nn.Identity()
    sci.kernel.data = torch.tensor(p["sci_kernel"], device=dev)
    cci.kernel.data = torch.tensor(p["cci_kernel"], device=dev)
    rbf.kernel.data = torch.tensor(p["rbf_kernel"], device=dev)
    x = torch.tensor(xn, device=dev)
    v = torch.tensor(vn, device=dev, requires_grad=True)
    c = cci(sci(x))
    rec = rbf(v, x)
    m = x[:, C:2 * C]
    loss = ((rec * m - x[:, :C] * m) ** 2).sum() / (m == 1.0).sum()     # pretrain_interp.py:169-175
    ((c * torch.tensor(gc, device=dev)).sum() + loss).backward()
    _check_groups("c1/cci_out", c, c64, C)
    _check("c1/rbf_out", rec, r64)
    _check("c1/rec_loss", loss, loss64)
    _check("c1/d_sci_kernel", sci.kernel.grad, dk64)
    _check("c1/d_cci_kernel", cci.kernel.grad, dK64)
    _check("c1/d_rbf_kernel", rbf.kernel.grad, dkr64)
    _check("c1/dv", v.grad, dv64)


def test_rbf_module_state_dict_dropin(golden):
    """The full RBF module (compress_fc included) loads the reference's state dict by key and
    reproduces its eval-mode output."""
    import deep_interpolation_clustering_b200 as dic
    g = golden("rbf_module")
    dev = torch.device("cuda:0")
    x = torch.tensor(g["x"], device=dev)
    C, T = x.shape[1] // 4, x.shape[2]
    rbf = dic.RBF(float(g["hours"]), int(g["R"]), g["interp"].shape[1], C, 0.2,
                  dic.basis_func_dict()["gaussian"], dev).to(dev)
    sd = {k[3:]: torch.tensor(v) for k, v in g.items() if k.startswith("sd.")}
    assert set(sd) == set(rbf.state_dict())
    rbf.load_state_dict(sd, strict=True)
    rbf.eval()
    with torch.no_grad():
        out = rbf(torch.tensor(g["interp"], device=dev), x)
    _check("rbf_module/out", out, g["out"], 1e-4)       # TF32-free cuBLAS vs CPU GEMM in compress_fc


def test_shape_and_device_errors():
    import deep_interpolation_clustering_b200 as dic
    dev = torch.device("cuda:0")
    sci = dic.SingleChannelInterp(8, 24.0, 6, 16, dev)
    with pytest.raises(RuntimeError):
        sci(torch.zeros(2, 24, 17, device=dev))          # timestamp mismatch, as the reference
    with pytest.raises(RuntimeError):
        sci(torch.zeros(2, 24, 16))                      # CPU tensor: no fallback
    with pytest.raises(TypeError):
        sci(torch.zeros(2, 24, 16, device=dev, dtype=torch.float64))


def test_batch_independence_and_constant_signal_large():
    """Size-independent properties at a BASELINE-config-2-shaped batch (T=256, R=96):
    (i) every encounter's output equals the output of that encounter alone (indexing at large B);
    (ii) a constant signal is reproduced by both filters (weights sum to 1)."""
    import deep_interpolation_clustering_b200 as dic
    from deep_interpolation_clustering_b200 import synth
    dev = torch.device("cuda:0")
    B, C, T, R, H = 65536, 6, 256, 96, 24.0
    x = synth.make_encounters_device(B, C, T, H, 5.0, seed=3, device=dev)
    sci = dic.SingleChannelInterp(R, H, C, T, dev)
    cci = dic.CrossChannelInterp(C, T, dev)
    with torch.no_grad():
        full = cci(sci(x))
        idx = torch.tensor([0, 1, 4097, 32768, B - 2, B - 1], device=dev)
        part = cci(sci(x[idx].contiguous()))
        assert torch.equal(full[idx], part)
        xc = x.clone()
        xc[:, :C] = 1.75 * xc[:, C:2 * C]
        s = sci(xc)
        y, y10 = s[:, :, :C], s[:, :, 2 * C:]
        assert torch.allclose(y, torch.full_like(y, 1.75), rtol=1e-5, atol=0)
        assert torch.allclose(y10, torch.full_like(y10, 1.75), rtol=1e-5, atol=0)
        assert torch.isfinite(s).all()


def test_unsorted_and_weighted_masks_vs_oracle():
    """Shuffled observation order and fractional mask weights (log m acts as a weight,
    interpolation_layer.py:59) follow the same closed form."""
    import deep_interpolation_clustering_b200 as dic
    from deep_interpolation_clustering_b200 import synth
    from oracle import interp_oracle
    dev = torch.device("cuda:0")
    B, C, T, R, H = 16, 6, 40, 24, 24.0
    xn = synth.make_adversarial_encounters(B, C, T, H, seed=9)
    rng = np.random.RandomState(10)
    w = rng.uniform(0.25, 1.0, size=(B, C, T)).astype(np.float32)
    xn[:, C:2 * C] *= w                                  # fractional weights on valid entries
    p = synth.make_interp_params(C, seed=1)
    rt = interp_oracle.linspace_grid(H, R)
    s64 = interp_oracle.sci_forward(xn.astype(np.float64), p["sci_kernel"].astype(np.float64), rt, C)
    vn = rng.normal(size=(B, C, R)).astype(np.float32)
    r64 = interp_oracle.rbf_forward(vn, xn.astype(np.float64), p["rbf_kernel"].astype(np.float64), rt, C)
    sci = dic.SingleChannelInterp(R, H, C, T, dev)
    rbf = dic.RBF(H, R, C, C, 0.0, dic.basis_func_dict()["gaussian"], dev)
    rbf.compress_fc = torch.nn.Identity()
    sci.kernel.data = torch.tensor(p["sci_kernel"], device=dev)
    rbf.kernel.data = torch.tensor(p["rbf_kernel"], device=dev)
    x = torch.tensor(xn, device=dev)
    _check_groups("weighted/sci", sci(x), s64, C)
    _check("weighted/rbf", rbf(torch.tensor(vn, device=dev), x), r64)
    gs = rng.normal(size=s64.shape).astype(np.float32)
    dk64 = interp_oracle.sci_backward(xn.astype(np.float64), p["sci_kernel"].astype(np.float64), rt, C, gs)
    (sci(x) * torch.tensor(gs, device=dev)).sum().backward()
    _check("weighted/d_sci_kernel", sci.kernel.grad, dk64)


@pytest.mark.parametrize("case", ["interp_c1", "interp_odd"])
def test_three_plane_input_is_bit_identical(golden, case):
    """x without its never-read hold-out plane - uploaded by dic_upload_encounters from the host
    (B,4C,T) tensor, or a strided slice x[:, :3C] of the dense device tensor - gives bit-identical
    forward outputs and gradients (interpolation_layer.py:26-30 only reads planes [0, 3C))."""
    import deep_interpolation_clustering_b200 as dic
    g = golden(case)
    dev = torch.device("cuda:0")
    C = g["x"].shape[1] // 4
    xh = torch.from_numpy(np.ascontiguousarray(g["x"])).pin_memory()
    x4 = xh.to(dev)
    x3 = dic.upload_encounters(xh, device=dev)
    torch.cuda.synchronize()
    assert x3.shape == (x4.shape[0], 3 * C, x4.shape[2])
    assert torch.equal(x3, x4[:, :3 * C])
    results = []
    for x in (x4, x3, x4[:, :3 * C]):
        sci, cci, rbf = _modules(g, dev)
        v = torch.tensor(g["v"], device=dev, requires_grad=True)
        c = cci(sci(x))
        r = rbf(v, x)
        (c * torch.tensor(g["g_cci"], device=dev)).sum().backward()
        (r * torch.tensor(g["g_rbf"], device=dev)).sum().backward()
        results.append([t.detach().clone() for t in (c, r, sci.kernel.grad, cci.kernel.grad, rbf.kernel.grad, v.grad)])
    for other in results[1:]:
        for a, b in zip(results[0], other):
            assert torch.equal(a, b) or (torch.isnan(a) == torch.isnan(b)).all() and torch.equal(
                torch.nan_to_num(a), torch.nan_to_num(b))


def test_empty_batch_and_extreme_observation_counts():
    """Edge cases: B = 0 (empty outputs, zero parameter gradients), one observation per vital
    (the softmax collapses onto it: y = y' = x_0 everywhere), every slot observed (n = T),
    a duplicated timestamp."""
    import deep_interpolation_clustering_b200 as dic
    from deep_interpolation_clustering_b200 import synth
    from oracle import interp_oracle
    dev = torch.device("cuda:0")
    C, T, R, H = 6, 32, 24, 24.0
    sci = dic.SingleChannelInterp(R, H, C, T, dev)
    cci = dic.CrossChannelInterp(C, T, dev)
    rbf = dic.RBF(H, R, C, C, 0.0, dic.basis_func_dict()["gaussian"], dev)
    rbf.compress_fc = torch.nn.Identity()
    # B = 0
    x0 = torch.zeros((0, 4 * C, T), device=dev)
    v0 = torch.zeros((0, C, R), device=dev, requires_grad=True)
    out0, rec0 = cci(sci(x0)), rbf(v0, x0)
    assert out0.shape == (0, R, 3 * C) and rec0.shape == (0, C, T)
    (out0.sum() + rec0.sum()).backward()
    for prm in (sci.kernel, cci.kernel, rbf.kernel):
        assert prm.grad is not None and float(prm.grad.abs().sum()) == 0.0
    # n_obs = 1 and n_obs = T, against the float64 oracle
    p = synth.make_interp_params(C, seed=1)
    rt = interp_oracle.linspace_grid(H, R)
    for tag, xn in (("one_obs", synth.make_encounters(8, C, T, H, seed=21)),
                    ("all_obs", synth.make_encounters(8, C, T, H, seed=22, min_obs=T))):
        if tag == "one_obs":
            xn[:, C:2 * C, 1:] = 0.0
            xn[:, :C, 1:] = 0.0
            xn[:, 2 * C:3 * C, 1:] = 0.0
        else:
            xn[:, 2 * C:3 * C, 5] = xn[:, 2 * C:3 * C, 4]          # duplicated timestamp
        sci.kernel.data = torch.tensor(p["sci_kernel"], device=dev)
        s64 = interp_oracle.sci_forward(xn.astype(np.float64), p["sci_kernel"].astype(np.float64), rt, C)
        got = sci(torch.tensor(xn, device=dev))
        _check_groups(f"edge_{tag}/sci", got, s64, C)
        if tag == "one_obs":
            y = got[:, :, :C].detach().cpu().numpy()
            assert np.array_equal(y, np.broadcast_to(xn[:, None, :C, 0], y.shape))


SHAPES = [  # (B, C, T, R): vitals != 6 (generic CCI path), T not a multiple of 4 (no TMA staging), R in every RPT /
            # chunk regime (1, 2, 3 points per lane; two chunks), tiny and ragged sizes
    (7, 1, 5, 3), (5, 3, 17, 33), (4, 6, 64, 64), (3, 8, 100, 100), (3, 6, 37, 192), (2, 4, 256, 97), (6, 2, 12, 48),
    (3, 10, 20, 16),
]


@pytest.mark.parametrize("B,C,T,R", SHAPES)
def test_shape_sweep_vs_oracle(B, C, T, R):
    """Forward and backward of SCI -> CCI and RBF against the float64 numpy oracle over shapes that take the
    non-default kernel paths."""
    import deep_interpolation_clustering_b200 as dic
    from deep_interpolation_clustering_b200 import synth
    from oracle import interp_oracle as O
    dev = torch.device("cuda:0")
    H = 24.0
    xn = synth.make_encounters(B, C, T, H, seed=B * 1000 + T)
    rng = np.random.RandomState(C * 100 + R)
    ks, kr = rng.uniform(size=C).astype(np.float32), rng.uniform(size=C).astype(np.float32)
    kc = (np.eye(C) + 0.1 * rng.normal(size=(C, C))).astype(np.float32)
    vn = rng.normal(size=(B, C, R)).astype(np.float32)
    gc = rng.normal(size=(B, R, 3 * C)).astype(np.float32)
    gr = rng.normal(size=(B, C, T)).astype(np.float32)
    rt = O.linspace_grid(H, R)
    x64 = xn.astype(np.float64)
    s64 = O.sci_forward(x64, ks.astype(np.float64), rt, C)
    c64 = O.cci_forward(s64, kc.astype(np.float64), C)
    r64 = O.rbf_forward(vn.astype(np.float64), x64, kr.astype(np.float64), rt, C)
    du64, dkc64 = O.cci_backward(s64, kc.astype(np.float64), C, gc.astype(np.float64))
    dks64 = O.sci_backward(x64, ks.astype(np.float64), rt, C, du64)
    dv64, dkr64 = O.rbf_backward(vn.astype(np.float64), x64, kr.astype(np.float64), rt, C, gr.astype(np.float64))

    sci = dic.SingleChannelInterp(R, H, C, T, dev)
    cci = dic.CrossChannelInterp(C, T, dev)
    rbf = dic.RBF(H, R, C, C, 0.0, dic.basis_func_dict()["gaussian"], dev)
    rbf.compress_fc = torch.nn.Identity()
    sci.kernel.data, cci.kernel.data, rbf.kernel.data = (torch.tensor(a, device=dev) for a in (ks, kc, kr))
    x = torch.tensor(xn, device=dev)
    v = torch.tensor(vn, device=dev, requires_grad=True)
    s = sci(x)
    c = cci(s)
    r = rbf(v, x)
    (c * torch.tensor(gc, device=dev)).sum().backward()
    (r * torch.tensor(gr, device=dev)).sum().backward()
    tag = f"sweep_B{B}C{C}T{T}R{R}"
    _check_groups(f"{tag}/sci", s, s64, C)
    _check_groups(f"{tag}/cci", c, c64, C)
    _check(f"{tag}/rbf", r, r64)
    _check(f"{tag}/d_cci_kernel", cci.kernel.grad, dkc64)
    # (coarse grids - R = 16: 1.6 h between grid points, shifts in the hundreds - are where the reference's own
    # float32 evaluation order is off by 3e-4 .. 8e-3 on d sci.kernel; the moment form stays within 1e-7)
    _check(f"{tag}/d_sci_kernel", sci.kernel.grad, dks64)
    _check(f"{tag}/dv", v.grad, dv64)
    _check(f"{tag}/d_rbf_kernel", rbf.kernel.grad, dkr64)


def test_allmasked_vital_contributes_zero_gradient(golden):
    """Backward with a finite upstream gradient: an all-masked vital (NaN / -inf outputs, the reference's gradient is
    NaN there) contributes exactly nothing to d kernel; the other vitals match the float64 oracle."""
    from deep_interpolation_clustering_b200 import functional as F_
    from oracle import interp_oracle as O
    g = golden("interp_allmasked")
    dev = torch.device("cuda:0")
    xn = g["x"]
    C, R, H = xn.shape[1] // 4, int(g["R"]), float(g["hours"])
    masked = [(b, c) for b in range(xn.shape[0]) for c in range(C) if xn[b, C + c].sum() == 0]
    assert masked, "fixture must contain an all-masked vital"
    k = torch.tensor(g["sci_kernel"], device=dev, requires_grad=True)
    rt = torch.linspace(0, H, R).to(dev)
    u = F_.sci(torch.tensor(xn, device=dev), k, rt)                     # planar (B, 3C, R)
    gu = np.random.RandomState(0).normal(size=tuple(u.shape)).astype(np.float32)
    u.backward(torch.tensor(gu, device=dev))
    assert torch.isfinite(k.grad).all()
    # oracle on a copy where the all-masked vitals get one dummy observation and a ZERO upstream gradient
    x2, g2 = xn.astype(np.float64).copy(), gu.astype(np.float64).copy()
    for b, c in masked:
        x2[b, C + c, 0] = 1.0
        g2[b, [c, C + c, 2 * C + c], :] = 0.0
    want = O.sci_backward(x2, g["sci_kernel"].astype(np.float64), O.linspace_grid(H, R), C, g2.transpose(0, 2, 1))
    _check("allmasked/d_sci_kernel", k.grad, want)


def test_observation_tensor_gradient_is_refused_loudly():
    """The operators are differentiable wrt their parameters (and the RBF grid values) only; an x that requires grad
    must raise instead of receiving a silent None (the reference's own mask-plane gradient is NaN, Appendix A.1)."""
    import deep_interpolation_clustering_b200 as dic
    from deep_interpolation_clustering_b200 import synth
    dev = torch.device("cuda:0")
    x = torch.tensor(synth.make_encounters(4, 6, 32, 24.0, seed=0), device=dev, requires_grad=True)
    sci = dic.SingleChannelInterp(24, 24.0, 6, 32, dev)
    rbf = dic.RBF(24.0, 24, 6, 6, 0.0, dic.basis_func_dict()["gaussian"], dev)
    rbf.compress_fc = torch.nn.Identity()
    with pytest.raises(RuntimeError, match="requires_grad"):
        sci(x)
    with pytest.raises(RuntimeError, match="requires_grad"):
        rbf(torch.zeros((4, 6, 24), device=dev), x)
    # a frozen kernel with a plain input still runs (no saved state, no backward)
    sci.kernel.requires_grad_(False)
    assert sci(x.detach()).shape == (4, 24, 18)


def test_config2_slice_of_4096_encounters_vs_the_staged_reference_in_float64():
    """"c2 B = 1M, T = 256, R = 96 (oracle checks a 4,096-row slice)".  The first 4,096 encounters of the c2
    generator through the kernels against the REFERENCE'S OWN modules (interpolation_layer.py, rbf.py as staged in
    baseline/_ref by oracle/make_ref.py - it travels with the tree) run in float64 on the host in chunks of 128:
    every output element and the three parameter gradients."""
    from deep_interpolation_clustering_b200 import synth
    import deep_interpolation_clustering_b200 as dic
    from oracle import make_ref
    ref = make_ref.import_reference()
    if ref is None:
        pytest.skip("no staged reference (python oracle/make_ref.py where /root/reference exists)")
    dev, cpu = torch.device("cuda:0"), torch.device("cpu")
    B, C, T, R, H = 4096, 6, 256, 96, 24.0
    xn = synth.make_encounters(B, C, T, H, seed=0)
    p = synth.make_interp_params(C, seed=1)
    rng = np.random.RandomState(2)
    vn = rng.normal(size=(B, C, R)).astype(np.float32)
    gc = rng.normal(size=(B, R, 3 * C)).astype(np.float32)
    gr = rng.normal(size=(B, C, T)).astype(np.float32)

    r_sci = ref.interpolation_layer.SingleChannelInterp(R, H, C, T, cpu).double()
    r_cci = ref.interpolation_layer.CrossChannelInterp(C, T, cpu).double()
    r_rbf = ref.rbf.RBF(H, R, C, C, 0.0, ref.rbf.basis_func_dict()["gaussian"], cpu)
    r_rbf.compress_fc = torch.nn.Identity()
    r_rbf = r_rbf.double()
    r_rbf.interp_t = r_rbf.interp_t.double()
    with torch.no_grad():
        r_sci.kernel.copy_(torch.from_numpy(p["sci_kernel"]).double())
        r_cci.kernel.copy_(torch.from_numpy(p["cci_kernel"]).double())
        r_rbf.kernel.copy_(torch.from_numpy(p["rbf_kernel"]).double())
    c64, rec64 = [], []
    torch.set_num_threads(max(1, (torch.get_num_threads())))
    for i in range(0, B, 128):
        xb = torch.from_numpy(xn[i:i + 128]).double()
        out = r_cci(r_sci(xb))
        rec = r_rbf(torch.from_numpy(vn[i:i + 128]).double(), xb)
        ((out * torch.from_numpy(gc[i:i + 128]).double()).sum() + (rec * torch.from_numpy(gr[i:i + 128]).double()).sum()).backward()
        c64.append(out.detach().numpy())
        rec64.append(rec.detach().numpy())
    c64, rec64 = np.concatenate(c64), np.concatenate(rec64)

    sci = dic.SingleChannelInterp(R, H, C, T, dev)
    cci = dic.CrossChannelInterp(C, T, dev)
    rbf = dic.RBF(H, R, C, C, 0.0, dic.basis_func_dict()["gaussian"], dev)
    rbf.compress_fc = torch.nn.Identity()
    sci.kernel.data = torch.tensor(p["sci_kernel"], device=dev)
    cci.kernel.data = torch.tensor(p["cci_kernel"], device=dev)
    rbf.kernel.data = torch.tensor(p["rbf_kernel"], device=dev)
    x = torch.tensor(xn, device=dev)
    c = cci(sci(x))
    rec = rbf(torch.tensor(vn, device=dev), x)
    ((c * torch.tensor(gc, device=dev)).sum() + (rec * torch.tensor(gr, device=dev)).sum()).backward()
    _check_groups("c2_slice4096/cci_out", c, c64, C)
    _check("c2_slice4096/rbf_out", rec, rec64)
    _check("c2_slice4096/d_sci_kernel", sci.kernel.grad, r_sci.kernel.grad.numpy())
    _check("c2_slice4096/d_cci_kernel", cci.kernel.grad, r_cci.kernel.grad.numpy())
    _check("c2_slice4096/d_rbf_kernel", rbf.kernel.grad, r_rbf.kernel.grad.numpy())


@pytest.mark.parametrize("case", ["interp_c2", "interp_odd", "interp_kernels", "interp_allmasked"])
def test_fused_cci_sci_backward_equals_the_two_kernel_chain(golden, case):
    """cci(sci(x)) through the modules folds the SCI backward into the CCI backward kernel (dic_cci_sci_bwd): the
    parameter gradients must equal those of the plain chain (dic_cci_bwd -> grad_u in HBM -> dic_sci_bwd) - same
    arithmetic on the same numbers, so to float32 rounding of the reduction order - and the float64 reference's."""
    from deep_interpolation_clustering_b200 import interpolation_layer as il
    g = golden(case)
    dev = torch.device("cuda:0")
    x = torch.tensor(g["x"], device=dev)
    gc = torch.tensor(g["g_cci"], device=dev)
    grads = {}
    for fused in (True, False):
        il.FUSE_SCI_CCI_BACKWARD = fused
        try:
            sci, cci, _ = _modules(g, dev)
            s = sci(x)
            c = cci(s)
            (c * gc).sum().backward()
            grads[fused] = (sci.kernel.grad.clone(), cci.kernel.grad.clone(), c.detach().clone())
        finally:
            il.FUSE_SCI_CCI_BACKWARD = True
    assert torch.allclose(grads[True][2], grads[False][2], rtol=0, atol=0, equal_nan=True)     # same forward kernel
    if case != "interp_allmasked":
        _check(f"{case}/fused_d_sci_kernel", grads[True][0], g["d_sci_kernel_f64"])
        _check(f"{case}/fused_d_cci_kernel", grads[True][1], g["d_cci_kernel_f64"])
    for a, b in zip(grads[True][:2], grads[False][:2]):
        fin = torch.isfinite(b)
        assert torch.equal(torch.isfinite(a), fin)
        assert torch.allclose(a[fin], b[fin], rtol=2e-6, atol=2e-6 * float(b[fin].abs().max() if fin.any() else 1.0))


def test_fusion_steps_aside_when_the_sci_output_gradient_is_observed(golden):
    g = golden("interp_c2")
    dev = torch.device("cuda:0")
    sci, cci, _ = _modules(g, dev)
    s = sci(torch.tensor(g["x"], device=dev))
    s.retain_grad()
    (cci(s) * torch.tensor(g["g_cci"], device=dev)).sum().backward()
    assert s.grad is not None and sci.kernel.grad is not None
    _check_groups("retain/d_sci_out", s.grad, g["d_sci_out_f64"], g["x"].shape[1] // 4)
