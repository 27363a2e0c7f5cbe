"""GPU parity of the DEC soft-assignment / target-distribution / KL step.

Truth = the reference (dec.py, clustering_interp.py:205-207) run in float64 on the golden
inputs; q and p must agree to 1e-6 absolute (BASELINE.json), gradients to rtol 1e-5 plus
the usual RMS-scaled floor.
"""
import numpy as np
import pytest
import torch

from gpu_util import ATOL_QP, RTOL_GRID, record, scale_atol

pytestmark = pytest.mark.gpu

CASES = ["dec_k4", "dec_k16", "dec_alpha2", "dec_k10"]


def _check(name, got, truth, rtol=RTOL_GRID, atol=None):
    truth = np.asarray(truth)
    return record(name, got.detach().cpu().numpy(), truth, rtol,
                  scale_atol(truth, rtol) if atol is None else atol)


@pytest.mark.parametrize("case", CASES)
def test_q_p_and_autograd(golden, case):
    import deep_interpolation_clustering_b200 as dic
    g = golden(case)
    dev = torch.device("cuda:0")
    K, D = g["mu"].shape
    ca = dic.ClusterAssignment(K, D, float(g["alpha"]), torch.tensor(g["mu"])).to(dev)
    z = torch.tensor(g["z"], device=dev, requires_grad=True)
    q = ca(z)
    p = dic.target_distribution(q).detach()
    _check(f"{case}/q", q, g["q_f64"], 0.0, ATOL_QP)
    _check(f"{case}/p", p, g["p_f64"], 0.0, ATOL_QP)
    kl = torch.nn.functional.kl_div(q.log(), p, reduction="batchmean")     # the trainer's own call
    kl.backward()
    _check(f"{case}/kl", kl, g["kl_f64"])
    _check(f"{case}/dz_kl", z.grad, g["dz_kl_f64"])
    _check(f"{case}/dmu_kl", ca.cluster_centers.grad, g["dmu_kl_f64"])
    z.grad = None
    ca.cluster_centers.grad = None
    (ca(z) * torch.tensor(g["gq"], device=dev)).sum().backward()
    _check(f"{case}/dz_gq", z.grad, g["dz_gq_f64"])
    _check(f"{case}/dmu_gq", ca.cluster_centers.grad, g["dmu_gq_f64"])


@pytest.mark.parametrize("case", CASES)
def test_fused_kl_step(golden, case):
    from deep_interpolation_clustering_b200 import functional as F_
    g = golden(case)
    dev = torch.device("cuda:0")
    z = torch.tensor(g["z"], device=dev)
    mu = torch.tensor(g["mu"], device=dev)
    out = F_.dec_kl_step(z, mu, float(g["alpha"]), weight=1.0)
    _check(f"{case}/fused_q", out["q"], g["q_f64"], 0.0, ATOL_QP)
    _check(f"{case}/fused_p", out["p"], g["p_f64"], 0.0, ATOL_QP)
    _check(f"{case}/fused_kl", out["kl"][0], g["kl_f64"])
    _check(f"{case}/fused_dz", out["grad_z"], g["dz_kl_f64"])
    _check(f"{case}/fused_dmu", out["grad_mu"], g["dmu_kl_f64"])
    assert np.array_equal(out["labels"].cpu().numpy(), np.argmax(g["q_f64"], axis=1))


def test_large_batch_vs_oracle_slices():
    """1M x 256 latents, K = 4 (BASELINE config 3 shape): q/p on slices against the float64
    oracle fed the device-computed global column sum; rows sum to 1; colsum == sum of q."""
    from deep_interpolation_clustering_b200 import functional as F_
    from deep_interpolation_clustering_b200 import synth
    from oracle import dec_oracle
    dev = torch.device("cuda:0")
    N, D, K = 1_000_000, 256, 4
    z, mu = synth.make_latents_device(N, D, K, seed=5, device=dev)
    out = F_.dec_kl_step(z, mu, 1.0, weight=10.0)
    q, p, f = out["q"], out["p"], out["colsum"]
    assert torch.allclose(q.sum(1), torch.ones(N, device=dev), atol=1e-5)
    assert torch.allclose(p.sum(1), torch.ones(N, device=dev), atol=1e-5)
    assert torch.allclose(f, q.double().sum(0), rtol=1e-9)
    fn = f.cpu().numpy()
    mun = mu.cpu().numpy().astype(np.float64)
    for lo in (0, 499_990, N - 4096):
        zs = z[lo:lo + 4096].cpu().numpy().astype(np.float64)
        q64 = dec_oracle.soft_assign(zs, mun, 1.0)
        p64 = dec_oracle.target_distribution(q64, colsum=fn)
        _check(f"dec1M/q@{lo}", q[lo:lo + 4096], q64, 0.0, ATOL_QP)
        _check(f"dec1M/p@{lo}", p[lo:lo + 4096], p64, 0.0, ATOL_QP)
        dz64, _ = dec_oracle.kl_backward_closed_form(zs, mun, p64, 1.0, batch=N, weight=10.0)
        _check(f"dec1M/dz@{lo}", out["grad_z"][lo:lo + 4096], dz64)
    assert np.array_equal(out["labels"][:4096].cpu().numpy(),
                          np.argmax(dec_oracle.soft_assign(z[:4096].cpu().numpy().astype(np.float64), mun), 1))


def test_errors():
    import deep_interpolation_clustering_b200 as dic
    dev = torch.device("cuda:0")
    ca = dic.ClusterAssignment(4, 64).to(dev)
    with pytest.raises(ValueError):
        ca(torch.zeros(8, 32, device=dev))               # embedding dimension mismatch
    with pytest.raises(RuntimeError):
        ca(torch.zeros(8, 64))                            # CPU tensor


def test_empty_batch():
    """B = 0: empty q / p, zero centre gradients, no launch."""
    import deep_interpolation_clustering_b200 as dic
    dev = torch.device("cuda:0")
    ca = dic.ClusterAssignment(4, 32, 1.0).to(dev)
    z = torch.zeros((0, 32), device=dev, requires_grad=True)
    q = ca(z)
    assert q.shape == (0, 4)
    p = dic.target_distribution(q)
    assert p.shape == (0, 4)
    q.sum().backward()
    assert float(ca.cluster_centers.grad.abs().sum()) == 0.0
