"""Seeded synthetic encounters, latents and feature matrices.

The reference is trained on private EHR data; every test and benchmark here uses
synthetic inputs shaped like what its pipeline produces:

* every (encounter, vital) has >= 1 observation and padding is left-packed
  (p0_data_process.py:65-67,88-92);
* timestamps are hours since admission in [0, H), 0 on padding (p0_data_process.py:55,67);
* values are min-max scaled to [-scale/2, scale/2], scale = 5 (dataloader.py:74-79),
  and multiplied by the padding mask before the model sees them
  (pretrain_trainer.py:136);
* the fourth plane (hold-out mask) is carried but never read by the hot path
  (interpolation_layer.py:26-30).

``numpy`` generators are used for parity tests (identical on every machine);
``*_device`` generators build BASELINE-size inputs directly in HBM.
"""
from __future__ import annotations

import numpy as np

__all__ = ["make_encounters", "make_adversarial_encounters", "make_interp_params",
           "make_latents", "make_blobs", "make_encounters_device", "make_latents_device"]


def make_encounters(B, C=6, T=64, hours=24.0, scale=5.0, seed=0, dtype=np.float32, min_obs=1):
    """(B, 4C, T) array: [value*mask | prefix mask | sorted times | hold-out]."""
    rng = np.random.RandomState(seed)
    n_obs = rng.randint(min_obs, T + 1, size=(B, C))
    idx = np.arange(T)[None, None, :]
    mask = (idx < n_obs[..., None]).astype(dtype)
    t = rng.uniform(0.0, hours, size=(B, C, T))
    # n_obs uniform draws per channel, sorted, left-packed; padding sorts to the end
    t = np.where(mask > 0, t, np.inf)
    t = np.sort(t, axis=-1)
    t = np.where(np.isfinite(t), t, 0.0).astype(dtype)
    val = rng.uniform(-scale / 2, scale / 2, size=(B, C, T)).astype(dtype) * mask
    return np.ascontiguousarray(np.concatenate([val, mask, t, mask], axis=1), dtype=dtype)


def make_adversarial_encounters(B, C=6, T=30, hours=24.0, scale=5.0, seed=0, dtype=np.float32):
    """Random (non-prefix) 0/1 masks with >= 1 observation per channel, unsorted times
    with N(0, 0.01) jitter that can dip below 0 (dataloader.py:207-208)."""
    rng = np.random.RandomState(seed)
    mask = (rng.uniform(size=(B, C, T)) < 0.5).astype(dtype)
    force = rng.randint(0, T, size=(B, C))
    bi, ci = np.meshgrid(np.arange(B), np.arange(C), indexing="ij")
    mask[bi, ci, force] = 1.0
    t = (rng.uniform(0.0, hours, size=(B, C, T)) + rng.normal(0, 0.01, size=(B, C, T)))
    t = (t * mask).astype(dtype)
    val = rng.uniform(-scale / 2, scale / 2, size=(B, C, T)).astype(dtype) * mask
    hold = (rng.uniform(size=(B, C, T)) < 0.8).astype(dtype)
    return np.ascontiguousarray(np.concatenate([val, mask, t, hold], axis=1), dtype=dtype)


def make_interp_params(C=6, seed=1, dtype=np.float32):
    """sci.kernel, rbf.kernel ~ U[0,1) (interpolation_layer.py:23, rbf.py:50);
    cci.kernel = I + 0.1 N(0,1) (identity is the init, :97; perturbed so the mix is exercised)."""
    rng = np.random.RandomState(seed)
    return dict(
        sci_kernel=rng.uniform(size=C).astype(dtype),
        cci_kernel=(np.eye(C) + 0.1 * rng.normal(size=(C, C))).astype(dtype),
        rbf_kernel=rng.uniform(size=C).astype(dtype),
    )


def make_latents(N, D=256, K=4, seed=0, spread=0.5, dtype=np.float32):
    """K Gaussian blobs: returns (z (N,D), centres (K,D))."""
    rng = np.random.RandomState(seed)
    centres = rng.normal(size=(K, D)).astype(dtype)
    lab = rng.randint(0, K, size=N)
    z = centres[lab] + spread * rng.normal(size=(N, D)).astype(dtype)
    return np.ascontiguousarray(z, dtype=dtype), centres


def make_blobs(N, D=64, n_blobs=5, seed=0, dtype=np.float32):
    """Feature matrix for the K-selection sweep: blobs sigma=1, centres 4*N(0,1)."""
    rng = np.random.RandomState(seed)
    centres = 4.0 * rng.normal(size=(n_blobs, D))
    lab = rng.randint(0, n_blobs, size=N)
    return np.ascontiguousarray(centres[lab] + rng.normal(size=(N, D)), dtype=dtype)


def make_encounters_device(B, C, T, hours, scale, seed, device, chunk=65536):
    """Same distribution as make_encounters, generated in HBM chunk by chunk (torch)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    x = torch.empty((B, 4 * C, T), dtype=torch.float32, device=device)
    idx = torch.arange(T, device=device)[None, None, :]
    for b0 in range(0, B, chunk):
        b1 = min(B, b0 + chunk)
        n = b1 - b0
        n_obs = torch.randint(1, T + 1, (n, C, 1), generator=g, device=device)
        mask = (idx < n_obs)
        t = torch.rand((n, C, T), generator=g, device=device) * hours
        t = torch.where(mask, t, torch.full_like(t, float("inf")))
        t, _ = torch.sort(t, dim=-1)
        t = torch.where(mask, t, torch.zeros_like(t))
        maskf = mask.to(torch.float32)
        val = (torch.rand((n, C, T), generator=g, device=device) - 0.5) * scale * maskf
        x[b0:b1, 0:C] = val
        x[b0:b1, C:2 * C] = maskf
        x[b0:b1, 2 * C:3 * C] = t
        x[b0:b1, 3 * C:4 * C] = maskf
    return x


def make_latents_device(N, D, K, seed, device, spread=0.5):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    centres = torch.randn((K, D), generator=g, device=device)
    lab = torch.randint(0, K, (N,), generator=g, device=device)
    z = centres[lab] + spread * torch.randn((N, D), generator=g, device=device)
    return z.contiguous(), centres
