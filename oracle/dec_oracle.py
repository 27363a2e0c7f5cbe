"""numpy restatement of the DEC soft-assignment / target-distribution / KL step.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Citations are into the upstream
repository: dec.py and clustering_interp.py.
"""
from __future__ import annotations

import numpy as np

__all__ = ["soft_assign", "soft_assign_backward", "target_distribution",
           "kl_div_batchmean", "kl_backward_closed_form"]


def _nu(z, mu, alpha):
    d2 = np.sum((z[:, None, :] - mu[None, :, :]) ** 2, axis=2)      # dec.py:56
    return 1.0 / (1.0 + d2 / alpha)                                 # dec.py:57


def soft_assign(z, mu, alpha=1.0):
    """ClusterAssignment.forward, dec.py:49-63.  z (B,D), mu (K,D) -> q (B,K)."""
    z = np.asarray(z)
    mu = np.asarray(mu, dtype=z.dtype)
    num = _nu(z, mu, z.dtype.type(alpha)) ** z.dtype.type((alpha + 1.0) / 2.0)   # dec.py:59-60
    return num / np.sum(num, axis=1, keepdims=True)                 # dec.py:61


def soft_assign_backward(z, mu, grad_q, alpha=1.0):
    """Gradient of soft_assign for an arbitrary upstream grad_q (B,K).

    With n = nu^((a+1)/2), q = n / sum_j n:
      dL/dn_ij  = (g_ij - sum_j' g_ij' q_ij') / S_i
      dn/dd2    = -((a+1)/(2a)) n nu
      c_ij      = -((a+1)/a) q_ij nu_ij (g_ij - <g_i, q_i>)
      dz_i = sum_j c_ij (z_i - mu_j),   dmu_j = -sum_i c_ij (z_i - mu_j)
    """
    z = np.asarray(z)
    mu = np.asarray(mu, dtype=z.dtype)
    g = np.asarray(grad_q, dtype=z.dtype)
    nu = _nu(z, mu, z.dtype.type(alpha))
    num = nu ** z.dtype.type((alpha + 1.0) / 2.0)
    q = num / np.sum(num, axis=1, keepdims=True)
    c = -((alpha + 1.0) / alpha) * q * nu * (g - np.sum(g * q, axis=1, keepdims=True))
    dz = np.sum(c, axis=1, keepdims=True) * z - c @ mu
    dmu = -(c.T.astype(np.float64) @ z.astype(np.float64)
            - np.sum(c, axis=0, dtype=np.float64)[:, None] * mu.astype(np.float64))
    return dz.astype(z.dtype), dmu.astype(z.dtype)


def target_distribution(q, colsum=None):
    """target_distribution, dec.py:66-76.

    ``colsum`` overrides f_j = sum_i q_ij (dec.py:73) so a shard of a larger
    batch can be normalised against the global column sum.
    """
    q = np.asarray(q)
    f = np.sum(q, axis=0) if colsum is None else np.asarray(colsum, dtype=q.dtype)
    w = q ** 2 / f                                                  # dec.py:73
    return (w.T / np.sum(w, axis=1)).T                              # dec.py:74


def kl_div_batchmean(p, q, batch=None):
    """Net.kl_loss, clustering_interp.py:205-207: F.kl_div(q.log(), p, 'batchmean').

    Terms with p == 0 contribute 0 (torch.xlogy semantics).
    """
    p = np.asarray(p)
    q = np.asarray(q)
    B = p.shape[0] if batch is None else batch
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.where(p > 0, p * (np.log(p) - np.log(q)), 0.0)
    return np.sum(t, dtype=np.float64) / B


def kl_backward_closed_form(z, mu, p, alpha=1.0, batch=None, weight=1.0):
    """Gradient of weight * KL(p || q(z, mu)) / B wrt z and mu with p detached
    (SURVEY Appendix A.4).  Equals soft_assign_backward with g = -weight p/(q B).
    """
    z = np.asarray(z)
    mu = np.asarray(mu, dtype=z.dtype)
    B = z.shape[0] if batch is None else batch
    nu = _nu(z, mu, z.dtype.type(alpha))
    q = soft_assign(z, mu, alpha)
    c = ((alpha + 1.0) / alpha) * (weight / B) * nu * (np.asarray(p, dtype=z.dtype) - q)
    dz = np.sum(c, axis=1, keepdims=True) * z - c @ mu
    dmu = -(c.T.astype(np.float64) @ z.astype(np.float64)
            - np.sum(c, axis=0, dtype=np.float64)[:, None] * mu.astype(np.float64))
    return dz.astype(z.dtype), dmu.astype(z.dtype)
