"""Drop-in mirror of the reference's ``interpolation_layer.py``.

Same class names, constructor signatures, parameter names (``kernel``), forward
signatures and tensor layouts as interpolation_layer.py:12-127, so that
``from interpolation_layer import SingleChannelInterp, CrossChannelInterp``
(pretrain_interp.py:11, clustering_interp.py:9) can resolve here unchanged - see
``dropin.install()``.  The arithmetic runs in the sm_100a kernels behind the C ABI.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import functional as F_


# cci(sci(x)): fold the SCI backward into the CCI backward kernel (functional.cci_after_sci).  The gradient of the SCI
# output then never exists as a tensor; set to False if code outside these modules asks autograd for it
# (torch.autograd.grad(..., inputs=sci_output)); retain_grad() / hooks on it are detected and keep the plain path.
FUSE_SCI_CCI_BACKWARD = True


class SingleChannelInterp(nn.Module):
    """Masked RBF interpolation of irregular observations onto a uniform grid.

    interpolation_layer.py:12-86.  ``forward(x)`` takes ``x (B, 4*d_dim, T)`` with planes
    [value | padding mask | timestamp (h) | hold-out (unused)] and returns ``(B, R, 3*d_dim)``
    = [low-pass y | log-intensity w | high-pass y'] as a permuted view of a (B, 3C, R)
    buffer (strides (3CR, 1, R)), exactly like the reference (:84-85).
    """

    def __init__(self, ref_points, hours_look_ahead, d_dim, timestamp, device, activation="sigmoid"):
        super().__init__()
        self.ref_points = ref_points
        self.hours_look_ahead = hours_look_ahead      # in hours
        self.activation = activation                   # stored, never used (:18)
        self.device = device
        self.timestamp = timestamp
        self.d_dim = d_dim
        self.kernel = nn.Parameter(torch.rand(self.d_dim, device=self.device), requires_grad=True)   # :23
        self._ref_t = None

    def _grid(self, device):
        # torch.linspace(0, H, R) as interpolation_layer.py:41; built on the CPU once so the
        # values are bit-identical to the CPU reference, then cached on the input's device.
        if self._ref_t is None or self._ref_t.device != device:
            self._ref_t = torch.linspace(0, self.hours_look_ahead, self.ref_points).to(device)
        return self._ref_t

    def forward(self, x):
        # (B, 3*d_dim, T) - the input without its never-read hold-out plane (:26-30) - works in the
        # reference too (it only slices [:d_dim], [d_dim:2d_dim], [2d_dim:3d_dim]) and is accepted here
        if x.dim() != 3 or x.shape[1] not in (3 * self.d_dim, 4 * self.d_dim):
            raise RuntimeError(f"expected input (B, {4 * self.d_dim}, T); got {tuple(x.shape)}")
        if x.shape[2] != self.timestamp:
            # the reference fails here with a broadcast error (:50-52)
            raise RuntimeError(f"The size of tensor a ({self.timestamp}) must match the size of tensor b "
                               f"({x.shape[2]}) at non-singleton dimension 1 (timestamp mismatch)")
        u = F_.sci(x, self.kernel, self._grid(x.device))
        out = u.permute(0, 2, 1)
        if hasattr(u, "_dic_sci"):
            out._dic_sci_planar = u          # lets a CrossChannelInterp fed with THIS tensor fuse the two backward passes
        return out


class CrossChannelInterp(nn.Module):
    """Cross-channel mixing of the SCI output.  interpolation_layer.py:89-127."""

    def __init__(self, d_dim, timestamp, device, activation="sigmoid"):
        super().__init__()
        self.d_dim = d_dim
        self.timestamp = timestamp
        self.device = device
        self.activation = activation
        gain = 1.0
        self.kernel = nn.Parameter(gain * torch.eye(self.d_dim, self.d_dim, device=self.device),
                                   requires_grad=True)                                           # :97

    def forward(self, x, reconstruction=False):      # `reconstruction` is unused upstream too (:99)
        if x.dim() != 3 or x.shape[2] != 3 * self.d_dim:
            raise RuntimeError(f"expected input (B, R, {3 * self.d_dim}); got {tuple(x.shape)}")
        self.output_dim = x.shape[1]
        # cci(sci(x)) - the chain of pretrain_interp.py:138-139 - runs ONE fused backward kernel.  Not when the caller looks
        # at the gradient of the SCI output itself (retain_grad / hooks on x): that tensor then takes the plain path.
        u = getattr(x, "_dic_sci_planar", None) if FUSE_SCI_CCI_BACKWARD else None
        if u is not None and (x.retains_grad or x._backward_hooks):
            u = None
        out = F_.cci_after_sci(u, self.kernel) if u is not None else None
        if out is None:
            out = F_.cci(x.permute(0, 2, 1), self.kernel)     # planar (B, 3C, R); a view for SCI's output
        return out.permute(0, 2, 1)
