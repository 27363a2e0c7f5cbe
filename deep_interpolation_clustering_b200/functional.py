"""Operators of the hot path: thin autograd wrappers over the C ABI (include/dic_b200.h).

Everything here takes and returns CUDA float32 tensors in the PLANAR layouts of the C
ABI; the nn.Module mirrors (interpolation_layer.py, rbf.py, dec.py) apply the
reference's permuted views on top.  No operator has a CPU path: a non-CUDA tensor
raises, a missing shared library raises.
"""
from __future__ import annotations

import torch

from . import _lib

__all__ = ["sci", "cci", "rbf_readout", "upload_encounters", "dec_soft_assign", "dec_target_distribution",
           "dec_assign", "dec_kl_from_colsum", "dec_kl_step", "colsum"]


def _require_cuda_f32(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} is on {t.device}: the B200 hot path has no CPU fallback "
                           "(move the module and its inputs to a CUDA device)")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32 (got {t.dtype})")


def _no_input_grad(x, name):
    """The interpolation operators produce no gradient with respect to the observation tensor: no caller of the
    reference needs one (the trainers never set requires_grad on the batch; the reference's own mask-plane gradient
    is NaN).  Fail loudly rather than hand autograd a silent None."""
    if x.requires_grad:
        raise RuntimeError(f"{name}.requires_grad is set, but the B200 interpolation operators are differentiable with "
                           f"respect to their parameters (and the RBF grid values) only; pass {name}.detach()")


def _planes(x, C, name):
    """x is (B, 4C, T) like the reference's input, or (B, 3C, T) when the never-read hold-out plane
    (interpolation_layer.py:26-30) was left on the host.  Rows of one encounter must be dense; the
    batch stride is free (a slice x[:, :3C] of a dense tensor is taken as it is).  Returns x (made
    dense only if its inner strides are not) and the batch stride in floats."""
    if x.dim() != 3 or x.shape[1] not in (3 * C, 4 * C):
        raise ValueError(f"{name} must be (B, {4 * C}, T) (or (B, {3 * C}, T) without the hold-out plane); "
                         f"got {tuple(x.shape)}")
    T = x.shape[2]
    if x.shape[0] > 0 and (x.stride(2) != 1 or x.stride(1) != T or x.stride(0) < 3 * C * T):
        x = x.contiguous()
    return x, (x.stride(0) if x.shape[0] > 1 else x.shape[1] * T)


def upload_encounters(x_host, out=None, stream=None, device=None):
    """Host -> device copy of x_host (B, 4C, T) [pin it for an asynchronous copy] that moves only the
    three live planes [value | mask | time] of every encounter (dic_upload_encounters: one strided
    DMA, 25 % fewer PCIe bytes).  Returns a (B, 3C, T) CUDA tensor that SingleChannelInterp / RBF
    accept in place of x."""
    if x_host.is_cuda or x_host.dtype != torch.float32 or x_host.dim() != 3 or x_host.shape[1] % 4:
        raise ValueError("x_host must be a float32 host tensor of shape (B, 4*d_dim, T)")
    if not x_host.is_contiguous():
        x_host = x_host.contiguous()
    B, P, T = x_host.shape
    C = P // 4
    if out is None:
        out = torch.empty((B, 3 * C, T), dtype=torch.float32, device=device or "cuda")
    if tuple(out.shape) != (B, 3 * C, T) or not out.is_contiguous() or not out.is_cuda:
        raise ValueError(f"out must be a dense CUDA tensor of shape {(B, 3 * C, T)}")
    st = (stream or torch.cuda.current_stream(out.device)).cuda_stream
    with torch.cuda.device(out.device):
        _lib.check(_lib.lib().dic_upload_encounters(_lib.ptr(out), x_host.data_ptr(), B, C, T, P, 3 * C, st),
                   "dic_upload_encounters")
    return out


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


class _SCI(torch.autograd.Function):
    """dic_sci_fwd / dic_sci_bwd.  x (B,4C|3C,T), kernel (C), ref_t (R) -> u (B,3C,R)."""

    @staticmethod
    def forward(ctx, x, kernel, ref_t, xs):
        B, _, T = x.shape
        C, R = kernel.numel(), ref_t.numel()
        ctx.xs = xs
        need_grad = ctx.needs_input_grad[1]      # grad mode is off inside forward; ask the ctx
        with torch.cuda.device(x.device):
            u = torch.empty((B, 3 * C, R), dtype=torch.float32, device=x.device)
            stats = torch.empty((B, 3 * C, R), dtype=torch.float32, device=x.device) if need_grad else None
            _lib.check(_lib.lib().dic_sci_fwd(_lib.ptr(x), _lib.ptr(kernel), _lib.ptr(ref_t), _lib.ptr(u),
                                              _lib.ptr(stats), B, C, T, R, xs, _lib.current_stream(x.device)),
                       "dic_sci_fwd")
        if need_grad:
            ctx.save_for_backward(x, kernel, ref_t, u, stats)
            ctx.mark_non_differentiable(stats)
            return u, stats
        return u, None

    @staticmethod
    def backward(ctx, grad_u, _grad_stats=None):
        x, kernel, ref_t, u, stats = ctx.saved_tensors
        B, _, T = x.shape
        C, R = kernel.numel(), ref_t.numel()
        grad_u = grad_u.contiguous()
        with torch.cuda.device(x.device):
            dk = torch.empty_like(kernel)
            ws = _ws(_lib.lib().dic_interp_bwd_workspace_bytes(B, C), x.device)
            _lib.check(_lib.lib().dic_sci_bwd(_lib.ptr(x), _lib.ptr(kernel), _lib.ptr(ref_t), _lib.ptr(u),
                                              _lib.ptr(stats), _lib.ptr(grad_u), _lib.ptr(dk), _lib.ptr(ws),
                                              B, C, T, R, ctx.xs, _lib.current_stream(x.device)), "dic_sci_bwd")
        return None, dk, None, None


def sci(x, kernel, ref_t):
    """SingleChannelInterp in planar layout: returns u (B, 3C, R) = rows [y | w | y']."""
    for t, n in ((x, "x"), (kernel, "kernel"), (ref_t, "ref_t")):
        _require_cuda_f32(t, n)
    _no_input_grad(x, "x")
    x, xs = _planes(x, kernel.numel(), "x")
    u, stats = _SCI.apply(x, kernel.contiguous(), ref_t.contiguous(), xs)
    if stats is not None:
        # what a CrossChannelInterp fed with this very tensor needs to fold the SCI backward into its own
        # (cci_after_sci): the saved moment rows and the parameter they carry the gradient of
        u._dic_sci = (stats, kernel)
    return u


class _CCI(torch.autograd.Function):
    """dic_cci_fwd / dic_cci_bwd on planar (B,3C,R) tensors."""

    @staticmethod
    def forward(ctx, u, kernel):
        B, C3, R = u.shape
        C = C3 // 3
        with torch.cuda.device(u.device):
            out = torch.empty_like(u)
            _lib.check(_lib.lib().dic_cci_fwd(_lib.ptr(u), _lib.ptr(kernel), _lib.ptr(out), B, C, R,
                                              _lib.current_stream(u.device)), "dic_cci_fwd")
        ctx.save_for_backward(u, kernel)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        u, kernel = ctx.saved_tensors
        B, C3, R = u.shape
        C = C3 // 3
        grad_out = grad_out.contiguous()
        with torch.cuda.device(u.device):
            gu = torch.empty_like(u)
            dk = torch.empty_like(kernel)
            ws = _ws(_lib.lib().dic_cci_bwd_workspace_bytes(B, C), u.device)
            _lib.check(_lib.lib().dic_cci_bwd(_lib.ptr(u), _lib.ptr(kernel), _lib.ptr(grad_out), _lib.ptr(gu),
                                              _lib.ptr(dk), _lib.ptr(ws), B, C, R,
                                              _lib.current_stream(u.device)), "dic_cci_bwd")
        return gu, dk


class _CCIAfterSCI(torch.autograd.Function):
    """cci(sci(x)) with ONE backward kernel (dic_cci_sci_bwd): the gradient of the SCI output never reaches HBM.
    u is the SCI output, detached (its own autograd node only sees gradients from other consumers of u)."""

    @staticmethod
    def forward(ctx, u, kernel, sci_kernel, stats):
        B, C3, R = u.shape
        C = C3 // 3
        with torch.cuda.device(u.device):
            out = torch.empty_like(u)
            _lib.check(_lib.lib().dic_cci_fwd(_lib.ptr(u), _lib.ptr(kernel), _lib.ptr(out), B, C, R,
                                              _lib.current_stream(u.device)), "dic_cci_fwd")
        ctx.save_for_backward(u, kernel, sci_kernel, stats)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        u, kernel, sci_kernel, stats = ctx.saved_tensors
        B, C3, R = u.shape
        C = C3 // 3
        grad_out = grad_out.contiguous()
        with torch.cuda.device(u.device):
            dk = torch.empty_like(kernel)
            dks = torch.empty_like(sci_kernel)
            ws = _ws(_lib.lib().dic_cci_sci_bwd_workspace_bytes(B, C), u.device)
            _lib.check(_lib.lib().dic_cci_sci_bwd(_lib.ptr(u), _lib.ptr(kernel), _lib.ptr(sci_kernel), _lib.ptr(stats),
                                                  _lib.ptr(grad_out), _lib.ptr(dk), _lib.ptr(dks), _lib.ptr(ws), B, C, R,
                                                  _lib.current_stream(u.device)), "dic_cci_sci_bwd")
        return None, dk, dks, None


def cci_after_sci(u, kernel):
    """CrossChannelInterp on a tensor that `sci` produced in this graph: same result and gradients as ``cci(u, kernel)``,
    with the SCI backward folded into the CCI backward kernel.  Returns None when `u` does not qualify (not an SCI
    output of this graph, gradients off, d_dim > 8): the caller then takes the plain ``cci``."""
    link = getattr(u, "_dic_sci", None)
    if link is None or not torch.is_grad_enabled() or kernel.shape[0] > 8:
        return None
    stats, sci_kernel = link
    if not (sci_kernel.requires_grad and u.is_contiguous() and u.shape[1] == 3 * kernel.shape[0]):
        return None
    return _CCIAfterSCI.apply(u.detach(), kernel.contiguous(), sci_kernel, stats)


def cci(u, kernel):
    """CrossChannelInterp in planar layout: u (B,3C,R) -> (B,3C,R) rows [z | exp(w) | y' - z]."""
    _require_cuda_f32(u, "x")
    _require_cuda_f32(kernel, "kernel")
    if u.dim() != 3 or u.shape[1] != 3 * kernel.shape[0] or kernel.shape[0] != kernel.shape[1]:
        raise ValueError(f"expected (B, 3*d_dim, R) with kernel (d_dim, d_dim); got {tuple(u.shape)}, "
                         f"{tuple(kernel.shape)}")
    return _CCI.apply(u.contiguous(), kernel.contiguous())


class _RBF(torch.autograd.Function):
    """dic_rbf_fwd / dic_rbf_bwd.  v (B,C,R), x (B,4C,T) -> rec (B,C,T)."""

    @staticmethod
    def forward(ctx, v, x, kernel, ref_t, xs):
        B, C, R = v.shape
        T = x.shape[2]
        ctx.xs = xs
        need_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[2]
        with torch.cuda.device(x.device):
            rec = torch.empty((B, C, T), dtype=torch.float32, device=x.device)
            inv = torch.empty_like(rec) if need_grad else None
            _lib.check(_lib.lib().dic_rbf_fwd(_lib.ptr(v), _lib.ptr(x), _lib.ptr(kernel), _lib.ptr(ref_t),
                                              _lib.ptr(rec), _lib.ptr(inv), B, C, T, R, xs,
                                              _lib.current_stream(x.device)), "dic_rbf_fwd")
        if need_grad:
            ctx.save_for_backward(v, x, kernel, ref_t, rec, inv)
        return rec

    @staticmethod
    def backward(ctx, grad_rec):
        v, x, kernel, ref_t, rec, inv = ctx.saved_tensors
        B, C, R = v.shape
        T = x.shape[2]
        grad_rec = grad_rec.contiguous()
        with torch.cuda.device(x.device):
            gv = torch.empty_like(v)
            dk = torch.empty_like(kernel)
            ws = _ws(_lib.lib().dic_interp_bwd_workspace_bytes(B, C), x.device)
            _lib.check(_lib.lib().dic_rbf_bwd(_lib.ptr(v), _lib.ptr(x), _lib.ptr(kernel), _lib.ptr(ref_t),
                                              _lib.ptr(rec), _lib.ptr(inv), _lib.ptr(grad_rec), _lib.ptr(gv),
                                              _lib.ptr(dk), _lib.ptr(ws), B, C, T, R, ctx.xs,
                                              _lib.current_stream(x.device)), "dic_rbf_bwd")
        return gv, None, dk, None, None


def rbf_readout(v, x, kernel, ref_t):
    """Gaussian RBF read-out of grid values v (B,C,R) at the observation times in x."""
    for t, n in ((v, "interp_data"), (x, "raw_input"), (kernel, "kernel"), (ref_t, "interp_t")):
        _require_cuda_f32(t, n)
    if x.dim() != 3 or v.dim() != 3 or x.shape[1] not in (3 * v.shape[1], 4 * v.shape[1]) or v.shape[0] != x.shape[0]:
        raise ValueError(f"expected v (B,C,R) and raw_input (B,4C,T); got {tuple(v.shape)}, {tuple(x.shape)}")
    if v.shape[2] != ref_t.numel():
        raise ValueError(f"interp_data has {v.shape[2]} grid points but ref_points is {ref_t.numel()}")
    _no_input_grad(x, "raw_input")
    x, xs = _planes(x, v.shape[1], "raw_input")
    return _RBF.apply(v.contiguous(), x, kernel.contiguous(), ref_t.contiguous(), xs)


class _DecQ(torch.autograd.Function):
    """dic_dec_q_fwd / dic_dec_q_bwd.  z (B,D), mu (K,D) -> q (B,K)."""

    @staticmethod
    def forward(ctx, z, mu, alpha):
        B, D = z.shape
        K = mu.shape[0]
        with torch.cuda.device(z.device):
            q = torch.empty((B, K), dtype=torch.float32, device=z.device)
            _lib.check(_lib.lib().dic_dec_q_fwd(_lib.ptr(z), _lib.ptr(mu), _lib.ptr(q), None, None, None, B, D,
                                                K, float(alpha), _lib.current_stream(z.device)),
                       "dic_dec_q_fwd")
        ctx.save_for_backward(z, mu)
        ctx.alpha = float(alpha)
        return q

    @staticmethod
    def backward(ctx, grad_q):
        z, mu = ctx.saved_tensors
        B, D = z.shape
        K = mu.shape[0]
        grad_q = grad_q.contiguous()
        with torch.cuda.device(z.device):
            gz = torch.empty_like(z) if ctx.needs_input_grad[0] else None
            gmu = torch.empty_like(mu)
            ws = _ws(_lib.lib().dic_dec_workspace_bytes(K, D), z.device)
            _lib.check(_lib.lib().dic_dec_q_bwd(_lib.ptr(z), _lib.ptr(mu), _lib.ptr(grad_q), _lib.ptr(gz),
                                                _lib.ptr(gmu), _lib.ptr(ws), B, D, K, ctx.alpha,
                                                _lib.current_stream(z.device)), "dic_dec_q_bwd")
        return gz, gmu, None


def dec_soft_assign(z, mu, alpha=1.0):
    """Student-t soft assignment q (B,K) of latents z (B,D) to centres mu (K,D)."""
    _require_cuda_f32(z, "batch")
    _require_cuda_f32(mu, "cluster_centers")
    if z.dim() != 2 or mu.dim() != 2 or z.shape[1] != mu.shape[1]:
        raise ValueError(f"expected batch (B,D) and centres (K,D); got {tuple(z.shape)}, {tuple(mu.shape)}")
    return _DecQ.apply(z.contiguous(), mu.contiguous(), alpha)


def colsum(a):
    """Deterministic float64 column sums of a (rows, cols) float32 CUDA matrix."""
    _require_cuda_f32(a, "a")
    a = a.contiguous()
    rows, cols = a.shape
    with torch.cuda.device(a.device):
        out = torch.empty(cols, dtype=torch.float64, device=a.device)
        ws = _ws(_lib.lib().dic_colsum_workspace_bytes(cols), a.device)
        _lib.check(_lib.lib().dic_colsum_f32(_lib.ptr(a), _lib.ptr(out), _lib.ptr(ws), rows, cols,
                                             _lib.current_stream(a.device)), "dic_colsum_f32")
    return out


@torch.no_grad()
def dec_target_distribution(q, colsum_f64=None):
    """p (B,K) from q (B,K); ``colsum_f64`` overrides f_j = sum_i q_ij (e.g. all-reduced)."""
    _require_cuda_f32(q, "batch")
    if q.dim() != 2:
        raise ValueError(f"expected (B, K); got {tuple(q.shape)}")
    q = q.contiguous()
    f = colsum(q) if colsum_f64 is None else colsum_f64.to(device=q.device, dtype=torch.float64).contiguous()
    B, K = q.shape
    with torch.cuda.device(q.device):
        p = torch.empty_like(q)
        _lib.check(_lib.lib().dic_dec_p(_lib.ptr(q), _lib.ptr(f), _lib.ptr(p), B, K,
                                        _lib.current_stream(q.device)), "dic_dec_p")
    return p


@torch.no_grad()
def dec_assign(z, mu, alpha=1.0):
    """Stage 1 of the fused DEC step: q (B,K), hard labels argmax_j q (B, int32) and the local
    column sum f_j = sum_i q_ij (K, float64) in one pass over z."""
    _require_cuda_f32(z, "batch")
    _require_cuda_f32(mu, "cluster_centers")
    z, mu = z.contiguous(), mu.contiguous()
    B, D = z.shape
    K = mu.shape[0]
    L = _lib.lib()
    with torch.cuda.device(z.device):
        q = torch.empty((B, K), dtype=torch.float32, device=z.device)
        labels = torch.empty(B, dtype=torch.int32, device=z.device)
        f = torch.empty(K, dtype=torch.float64, device=z.device)
        ws = _ws(L.dic_dec_workspace_bytes(K, D), z.device)
        _lib.check(L.dic_dec_q_fwd(_lib.ptr(z), _lib.ptr(mu), _lib.ptr(q), _lib.ptr(labels), _lib.ptr(f),
                                   _lib.ptr(ws), B, D, K, float(alpha), _lib.current_stream(z.device)),
                   "dic_dec_q_fwd")
    return dict(q=q, labels=labels, colsum=f)


@torch.no_grad()
def dec_kl_from_colsum(z, mu, colsum_f64, alpha=1.0, weight=1.0, batch=None, want_p=True, want_grad_z=True):
    """Stage 2: p = target(q) with the given (global) column sum, the 'batchmean' KL times
    ``weight`` and its closed-form gradients wrt z and mu."""
    _require_cuda_f32(z, "batch")
    _require_cuda_f32(mu, "cluster_centers")
    z, mu = z.contiguous(), mu.contiguous()
    B, D = z.shape
    K = mu.shape[0]
    Bg = B if batch is None else int(batch)
    L = _lib.lib()
    f = colsum_f64.to(device=z.device, dtype=torch.float64).contiguous()
    with torch.cuda.device(z.device):
        ws = _ws(L.dic_dec_workspace_bytes(K, D), z.device)
        p = torch.empty((B, K), dtype=torch.float32, device=z.device) if want_p else None
        kl = torch.empty(1, dtype=torch.float64, device=z.device)
        gz = torch.empty_like(z) if want_grad_z else None
        gmu = torch.empty_like(mu)
        _lib.check(L.dic_dec_kl_fwd_bwd(_lib.ptr(z), _lib.ptr(mu), _lib.ptr(f), _lib.ptr(p), _lib.ptr(kl),
                                        _lib.ptr(gz), _lib.ptr(gmu), _lib.ptr(ws), B, D, K, float(alpha),
                                        float(weight) / Bg, _lib.current_stream(z.device)), "dic_dec_kl_fwd_bwd")
    return dict(p=p, kl=kl * (float(weight) / Bg), grad_z=gz, grad_mu=gmu)


def dec_kl_step(z, mu, alpha=1.0, weight=1.0, batch=None, colsum_f64=None, want_p=True, want_grad_z=True):
    """Fused DEC step (no autograd graph): q + column sum -> p, KL, closed-form gradients.

    Returns dict(q, labels, colsum, p, kl, grad_z, grad_mu) where ``kl`` is the 'batchmean' KL
    (sum / batch) times ``weight`` and the gradients are of that value.  ``colsum_f64`` lets a
    caller supply the all-reduced column sum of a sharded batch; ``batch`` is the global batch
    size for the 1/B factor.
    """
    out = dec_assign(z, mu, alpha)
    f = out["colsum"] if colsum_f64 is None else colsum_f64
    out.update(dec_kl_from_colsum(z, mu, f, alpha, weight, batch, want_p, want_grad_z))
    out["colsum"] = f
    return out
