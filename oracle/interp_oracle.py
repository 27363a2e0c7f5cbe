"""numpy restatement of the interpolation network's three operators.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Each function names the
reference lines it restates; all paths are into the upstream repository.

Layouts (identical to the reference):
  x           (B, 4C, T)   planes [value | padding mask | time (h) | hold-out]
  sci output  (B, R, 3C)   channels [low-pass y | log-intensity w | high-pass y']
  cci output  (B, R, 3C)   channels [mixed z | intensity exp(w) | y' - z]
  rbf input   v (B, C, R)  (= compress_fc output, the custom-kernel boundary)
  rbf output  (B, C, T)

All functions are dtype-generic: they compute in the dtype of ``x`` (float64
gives the "truth" the CUDA kernels are judged against, float32 mimics the
reference's own rounding regime).
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "softplus", "linspace_grid", "sci_forward", "sci_backward", "cci_forward",
    "cci_backward", "rbf_forward", "rbf_backward", "rec_loss", "rec_loss_grad",
]

_CHUNK = 64  # encounters per (B,C,T,R) temporary


def softplus(k):
    """log(1 + exp(k)) exactly as interpolation_layer.py:51 / rbf.py:78 (no threshold)."""
    return np.log(1.0 + np.exp(k))


def _sigmoid(k):
    return 1.0 / (1.0 + np.exp(-k))


def linspace_grid(hours, ref_points, dtype=np.float32):
    """Reference grid r_j of interpolation_layer.py:41 / rbf.py:43.

    torch.linspace computes in the output dtype with a symmetric formula
    (start + step*i on the lower half, end - step*(n-1-i) on the upper half);
    restated here so the oracle does not need torch.
    """
    dtype = np.dtype(dtype).type
    n = int(ref_points)
    if n == 1:
        return np.array([0.0], dtype=dtype)
    step = dtype(dtype(hours) / dtype(n - 1))
    i = np.arange(n)
    if dtype is np.float32:
        # torch's CPU kernel fuses the multiply-add (single rounding): emulate in float64,
        # where float32 x small-int products are exact
        lo = (np.float64(step) * i).astype(dtype)
        hi = (np.float64(hours) - np.float64(step) * (n - 1 - i)).astype(dtype)
    else:
        lo = step * i.astype(dtype)
        hi = dtype(hours) - step * (n - 1 - i).astype(dtype)
    return np.where(i < n // 2, lo, hi).astype(dtype)


def _planes(x, C):
    return x[:, :C, :], x[:, C:2 * C, :], x[:, 2 * C:3 * C, :]


def _lse(s, axis):
    """logsumexp with torch semantics: all -inf -> -inf (no NaN)."""
    mx = np.max(s, axis=axis, keepdims=True)
    safe = np.where(np.isfinite(mx), mx, 0.0)
    with np.errstate(divide="ignore"):
        out = np.log(np.sum(np.exp(s - safe), axis=axis, keepdims=True)) + safe
    return np.squeeze(out, axis=axis)


def _sci_core(xv, m, d, alpha, ref_t):
    """One chunk of SCI: returns y, w, y', and the two softmax tensors p, p'."""
    norm = (d[..., None] - ref_t) ** 2                       # (b,C,T,R)  :46-49
    with np.errstate(divide="ignore"):
        logm = np.log(m)[..., None]                          # -inf at padding  :59
    a = alpha[None, :, None, None]
    s = -a * norm + logm
    w = _lse(s, 2)                                           # :59
    with np.errstate(invalid="ignore"):
        p = np.exp(s - w[:, :, None, :])                     # :62-63
        y = np.sum(p * xv[..., None], axis=2)                # :64
        s10 = -10.0 * a * norm + logm                        # kappa = 10  :80
        w10 = _lse(s10, 2)
        p10 = np.exp(s10 - w10[:, :, None, :])               # :81-82
        y10 = np.sum(p10 * xv[..., None], axis=2)            # :83
    return y, w, y10, p, p10, norm


def sci_forward(x, kernel, ref_t, C):
    """SingleChannelInterp.forward, interpolation_layer.py:31-86.

    Returns the (B, R, 3C) array [y | w | y'] (a C-contiguous copy of the
    reference's permuted view).
    """
    x = np.asarray(x)
    dt = x.dtype
    ref_t = np.asarray(ref_t, dtype=dt)
    alpha = softplus(np.asarray(kernel, dtype=dt))
    B, R = x.shape[0], ref_t.shape[0]
    out = np.empty((B, R, 3 * C), dtype=dt)
    for b0 in range(0, B, _CHUNK):
        xv, m, d = _planes(x[b0:b0 + _CHUNK], C)
        y, w, y10, _, _, _ = _sci_core(xv, m, d, alpha, ref_t)
        out[b0:b0 + _CHUNK] = np.concatenate([y, w, y10], axis=1).transpose(0, 2, 1)  # :84-85
    return out


def sci_backward(x, kernel, ref_t, C, grad_out):
    """Closed-form gradient of sci_forward wrt ``kernel``.

    ``grad_out`` has the forward output's shape (B, R, 3C).  Input gradients are
    not produced: no caller of the reference needs them and the reference's own
    mask-plane gradient is NaN.
    """
    x = np.asarray(x)
    dt = x.dtype
    ref_t = np.asarray(ref_t, dtype=dt)
    k = np.asarray(kernel, dtype=dt)
    alpha = softplus(k)
    g = np.asarray(grad_out, dtype=dt).transpose(0, 2, 1)    # (B,3C,R)
    dalpha = np.zeros(C, dtype=np.float64)
    for b0 in range(0, x.shape[0], _CHUNK):
        xv, m, d = _planes(x[b0:b0 + _CHUNK], C)
        gy, gw, gy10 = (g[b0:b0 + _CHUNK, i * C:(i + 1) * C] for i in range(3))
        y, w, y10, p, p10, norm = _sci_core(xv, m, d, alpha, ref_t)
        ds = p * ((xv[..., None] - y[:, :, None, :]) * gy[:, :, None, :] + gw[:, :, None, :])
        ds10 = p10 * (xv[..., None] - y10[:, :, None, :]) * gy10[:, :, None, :]
        dalpha += np.sum(-norm * (ds + 10.0 * ds10), axis=(0, 2, 3), dtype=np.float64)
    return (dalpha * _sigmoid(k.astype(np.float64))).astype(dt)


def _cci_split(u, C):
    u = np.asarray(u)
    return u[..., :C], u[..., C:2 * C], u[..., 2 * C:3 * C]   # each (B,R,C)


def cci_forward(u, kernel, C):
    """CrossChannelInterp.forward, interpolation_layer.py:99-127.

    ``u`` is the (B, R, 3C) SCI output; returns (B, R, 3C) = [z | exp(w) | y' - z].
    """
    y, w, y10 = _cci_split(u, C)
    K = np.asarray(kernel, dtype=y.dtype)
    what = np.exp(w - _lse(w, 2)[..., None])                 # softmax over channels  :107-110
    mean = np.mean(y, axis=1, keepdims=True)                 # over reference points  :111-112
    z = (what * (y - mean)) @ K + mean                       # :113
    return np.concatenate([z, np.exp(w), y10 - z], axis=2)   # :104,:122-126


def cci_backward(u, kernel, C, grad_out):
    """Closed-form gradients of cci_forward.

    Returns (grad_u (B,R,3C), grad_kernel (C,C)).
    """
    y, w, y10 = _cci_split(u, C)
    dt = y.dtype
    K = np.asarray(kernel, dtype=dt)
    gz, gi, gt = _cci_split(np.asarray(grad_out, dtype=dt), C)
    R = y.shape[1]
    what = np.exp(w - _lse(w, 2)[..., None])
    mean = np.mean(y, axis=1, keepdims=True)
    yc = y - mean
    gzt = gz - gt                                            # z enters [z] and [y' - z]
    a = what * yc                                            # (B,R,C) left operand of @K
    dK = np.einsum("brc,brd->cd", a.astype(np.float64), gzt.astype(np.float64)).astype(dt)
    uu = gzt @ K.T                                           # (B,R,C): sum_c' K[c,c'] gzt[c']
    dy = what * uu - np.mean(what * uu, axis=1, keepdims=True) + np.mean(gzt, axis=1, keepdims=True)
    t = yc * uu
    dw = what * (t - np.sum(what * t, axis=2, keepdims=True)) + np.exp(w) * gi
    return np.concatenate([dy, dw, gt], axis=2), dK
    # NB: R appears only through the means above.


def _rbf_core(m, d, beta, ref_t):
    norm = (d[..., None] - ref_t) ** 2                       # rbf.py:75-76 (sqrt then square)
    phi = np.exp(-beta[None, :, None, None] * norm) * m[..., None]   # :95-96, gaussian :129-131
    return phi, norm


def rbf_forward(v, x, kernel, ref_t, C):
    """RBF.forward after compress_fc, rbf.py:57-108 (gaussian basis, rbf.py:129-131).

    ``v`` (B, C, R) is compress_fc's output permuted as in rbf.py:101-103.
    Returns y_hat (B, C, T).
    """
    x = np.asarray(x)
    dt = x.dtype
    ref_t = np.asarray(ref_t, dtype=dt)
    beta = softplus(np.asarray(kernel, dtype=dt))            # :78
    v = np.asarray(v, dtype=dt)
    out = np.empty((x.shape[0], C, x.shape[2]), dtype=dt)
    for b0 in range(0, x.shape[0], _CHUNK):
        _, m, d = _planes(x[b0:b0 + _CHUNK], C)
        phi, _ = _rbf_core(m, d, beta, ref_t)
        nrm = np.sum(phi, axis=-1)                           # :97
        num = np.sum(phi * v[b0:b0 + _CHUNK, :, None, :], axis=-1)   # :105-106
        out[b0:b0 + _CHUNK] = num / (nrm + dt.type(1e-10)) * m       # :107
    return out


def rbf_backward(v, x, kernel, ref_t, C, grad_out):
    """Closed-form gradients of rbf_forward.

    Returns (grad_v (B,C,R), grad_kernel (C,)).
    """
    x = np.asarray(x)
    dt = x.dtype
    ref_t = np.asarray(ref_t, dtype=dt)
    k = np.asarray(kernel, dtype=dt)
    beta = softplus(k)
    v = np.asarray(v, dtype=dt)
    g = np.asarray(grad_out, dtype=dt)
    dv = np.empty_like(v)
    dbeta = np.zeros(C, dtype=np.float64)
    for b0 in range(0, x.shape[0], _CHUNK):
        _, m, d = _planes(x[b0:b0 + _CHUNK], C)
        vb = v[b0:b0 + _CHUNK]
        phi, norm = _rbf_core(m, d, beta, ref_t)
        den = np.sum(phi, axis=-1) + dt.type(1e-10)
        S = np.sum(phi * vb[:, :, None, :], axis=-1) / den
        a = m * g[b0:b0 + _CHUNK] / den                      # (b,C,T)
        dv[b0:b0 + _CHUNK] = np.sum(a[..., None] * phi, axis=2)
        dphi = a[..., None] * (vb[:, :, None, :] - S[..., None])
        dbeta += np.sum(-norm * phi * dphi, axis=(0, 2, 3), dtype=np.float64)
    return dv, (dbeta * _sigmoid(k.astype(np.float64))).astype(dt)


def rec_loss(x, rec, C):
    """Masked reconstruction MSE, pretrain_interp.py:169-175 (drives RBF grads)."""
    ob, m, _ = _planes(np.asarray(x), C)
    n = np.sum(m == 1.0)
    return np.sum((rec * m - ob * m) ** 2) / n


def rec_loss_grad(x, rec, C):
    """d rec_loss / d rec."""
    ob, m, _ = _planes(np.asarray(x), C)
    n = np.sum(m == 1.0)
    return 2.0 * (rec * m - ob * m) * m / n
