"""torch-CPU port of the reference's hot-path modules (autograd, op by op).

TEST INFRASTRUCTURE / CPU BASELINE ARM (see oracle/__init__.py).  The reference cannot
travel to the GPU box (``/root/reference`` is absent there), so the CPU arm that
``bench.py`` times there is this port: the same ATen-level algorithm the reference runs -
materialised (B, C, T, R) tensors, logsumexp + re-exponentiation for both filters, autograd
for the backward - restated from the closed forms, not copied.  It is pinned against the
reference's own outputs in tests/test_oracle_golden.py::test_ref_port_*.

    SingleChannelInterp   interpolation_layer.py:31-86
    CrossChannelInterp    interpolation_layer.py:99-127
    RBFReadout            rbf.py:57-108 (+ compress_fc, rbf.py:111-125)
    soft_assign / target_distribution / kl   dec.py:49-76, clustering_interp.py:205-207
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


def _softplus(k):
    return torch.log(1 + torch.exp(k))        # plain formula, interpolation_layer.py:51


class SingleChannelInterp(nn.Module):
    def __init__(self, ref_points, hours, d_dim):
        super().__init__()
        self.R, self.H, self.C = ref_points, hours, d_dim
        self.kernel = nn.Parameter(torch.rand(d_dim))

    def forward(self, x):
        C = self.C
        val, m, d = x[:, :C, :, None], x[:, C:2 * C, :, None], x[:, 2 * C:3 * C, :, None]
        grid = torch.linspace(0, self.H, self.R, dtype=x.dtype)
        sq = (d - grid) ** 2                                        # (B,C,T,R)
        alpha = _softplus(self.kernel)[None, :, None, None]
        logm = torch.log(m)
        outs = []
        for kappa in (1.0, 10.0):
            s = -kappa * alpha * sq + logm
            lse = torch.logsumexp(s, dim=2)
            outs.append((lse, (torch.exp(s - lse[:, :, None, :]) * val).sum(dim=2)))
        (w, y), (_, y10) = outs
        return torch.cat([y, w, y10], dim=1).permute(0, 2, 1)       # (B,R,3C)


class CrossChannelInterp(nn.Module):
    def __init__(self, d_dim):
        super().__init__()
        self.C = d_dim
        self.kernel = nn.Parameter(torch.eye(d_dim))

    def forward(self, u):
        C = self.C
        y, w, y10 = u[..., :C], u[..., C:2 * C], u[..., 2 * C:]
        what = torch.softmax(w, dim=2)
        mean = y.mean(dim=1, keepdim=True)
        z = (what * (y - mean)) @ self.kernel + mean
        return torch.cat([z, torch.exp(w), y10 - z], dim=2)


class RBFReadout(nn.Module):
    """Gaussian read-out; ``compress`` is applied per grid point when given (else v is the input)."""

    def __init__(self, ref_points, hours, d_dim, compress=None):
        super().__init__()
        self.R, self.H, self.C = ref_points, hours, d_dim
        self.kernel = nn.Parameter(torch.rand(d_dim))
        self.compress = compress

    def forward(self, v, x):
        C = self.C
        m, d = x[:, C:2 * C, :], x[:, 2 * C:3 * C, :]
        if self.compress is not None:
            B, Din, R = v.shape
            v = self.compress(v.permute(0, 2, 1).reshape(B * R, Din)).reshape(B, R, C).permute(0, 2, 1)
        grid = torch.linspace(0, self.H, self.R, dtype=x.dtype)
        beta = _softplus(self.kernel)[None, :, None, None]
        phi = torch.exp(-beta * (d[..., None] - grid) ** 2) * m[..., None]
        return (phi * v[:, :, None, :]).sum(-1) / (phi.sum(-1) + 1e-10) * m


def soft_assign(z, mu, alpha=1.0):
    d2 = ((z[:, None, :] - mu[None]) ** 2).sum(2)
    num = (1.0 / (1.0 + d2 / alpha)) ** ((alpha + 1.0) / 2)
    return num / num.sum(1, keepdim=True)


def target_distribution(q):
    w = q ** 2 / q.sum(0)
    return (w.t() / w.sum(1)).t()


def kl(p, q):
    return F.kl_div(q.log(), p, reduction="batchmean")


def masked_mse(x, rec, C):
    ob, m = x[:, :C], x[:, C:2 * C]
    return ((rec * m - ob * m) ** 2).sum() / (m == 1.0).sum()


def interp_step(sci, cci, rbf, x, v, g_cci):
    """One fwd+bwd of the interpolation path as bench.py times it: SCI -> CCI (driven by a
    fixed random projection g_cci so every output channel gets gradient) and the RBF read-out
    driven by the masked reconstruction MSE (pretrain_interp.py:169-175)."""
    out = cci(sci(x))
    rec = rbf(v, x)
    loss = (out * g_cci).sum() / x.shape[0] + masked_mse(x, rec, sci.C)
    loss.backward()
    return loss.detach()


def dec_step(z, mu, alpha=1.0):
    q = soft_assign(z, mu, alpha)
    p = target_distribution(q).detach()
    loss = kl(p, q)
    loss.backward()
    return loss.detach(), q.detach(), p
