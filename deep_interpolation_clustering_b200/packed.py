"""Ragged (packed) encounters: the host-side format that ships only what the left-packed rows of the
pipeline hold (p0_data_process.py:44-67) instead of the dense, half-padding planes the trainer moves
(pretrain_trainer.py:132-136).

    pe = PackedEncounters.from_dense(x_host)          # once per data set: (B, 4C, T) float32 host tensor
    st = PackedStaging.for_chunks(pe, 32768)          # device staging for one chunk
    x  = st.upload(pe, b0, b1, out=xbuf)              # H2D of ~1/3 of the bytes + device-side expansion
    y  = sci(x); rec = rbf(v, x)                      # (B, 3C, T) dense planes, bit-identical to the dense upload

Layout: see include/dic_b200.h ("ragged (packed) encounters").  Only 0/1 prefix masks can be packed; anything
else raises ``ValueError`` and goes through ``upload_encounters`` (the dense path).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib

__all__ = ["PackedEncounters", "PackedStaging"]


class PackedEncounters:
    """Host-resident packed encounters (pinned when CUDA is available)."""

    def __init__(self, n_obs, enc_off, packed, C, T, all_sorted):
        self.n_obs, self.enc_off, self.packed = n_obs, enc_off, packed
        self.C, self.T, self.all_sorted = int(C), int(T), bool(all_sorted)

    @property
    def B(self):
        return self.n_obs.shape[0]

    def __len__(self):
        return self.B

    @classmethod
    def from_dense(cls, x_host, pin=None):
        """Pack x_host (B, 4C or 3C, T) float32 on the host (dic_pack_encounters_host: plain C, no device work)."""
        if isinstance(x_host, np.ndarray):
            x_host = torch.from_numpy(np.ascontiguousarray(x_host, dtype=np.float32))
        if x_host.is_cuda or x_host.dtype != torch.float32 or x_host.dim() != 3:
            raise ValueError("x_host must be a float32 host tensor of shape (B, 4*d_dim, T)")
        x_host = x_host.contiguous()
        B, P, T = x_host.shape
        if P % 4 == 0:
            C = P // 4
        elif P % 3 == 0:
            C = P // 3
        else:
            raise ValueError(f"x_host has {P} planes: expected 4*d_dim (or 3*d_dim without the hold-out plane)")
        pin = torch.cuda.is_available() if pin is None else pin
        L = _lib.lib()
        n_obs = torch.empty((B, C), dtype=torch.int32)
        enc_off = torch.empty(B + 1, dtype=torch.int64)
        srt = ctypes.c_int(0)
        total = L.dic_pack_encounters_host(x_host.data_ptr(), B, C, T, P, n_obs.data_ptr(), enc_off.data_ptr(),
                                           None, 0, ctypes.byref(srt))
        if total < 0:
            _lib.check(int(total), "dic_pack_encounters_host")
        packed = torch.empty(max(int(total), 4), dtype=torch.float32)
        if pin:
            n_obs, enc_off, packed = n_obs.pin_memory(), enc_off.pin_memory(), packed.pin_memory()
        total2 = L.dic_pack_encounters_host(x_host.data_ptr(), B, C, T, P, n_obs.data_ptr(), enc_off.data_ptr(),
                                            packed.data_ptr(), packed.numel(), ctypes.byref(srt))
        if total2 < 0:
            _lib.check(int(total2), "dic_pack_encounters_host")
        return cls(n_obs, enc_off, packed, C, T, srt.value)

    def floats(self, b0=0, b1=None):
        b1 = self.B if b1 is None else b1
        return int(self.enc_off[b1]) - int(self.enc_off[b0])

    def nbytes(self, b0=0, b1=None):
        """Bytes dic_upload_encounters_packed moves for encounters [b0, b1)."""
        b1 = self.B if b1 is None else b1
        return 4 * self.floats(b0, b1) + 4 * (b1 - b0) * self.C + 8 * (b1 - b0 + 1)

    def max_chunk_floats(self, chunk):
        """Largest packed size of any aligned chunk of `chunk` encounters (sizes a PackedStaging)."""
        off = self.enc_off.numpy()
        starts = np.arange(0, self.B, chunk)
        ends = np.minimum(starts + chunk, self.B)
        return int((off[ends] - off[starts]).max()) if self.B else 0

    def to_dense(self):
        """Host-side expansion back to (B, 3C, T) (tests / debugging; numpy)."""
        B, C, T = self.B, self.C, self.T
        out = np.zeros((B, 3 * C, T), dtype=np.float32)
        n, off, pk = self.n_obs.numpy(), self.enc_off.numpy(), self.packed.numpy()
        for b in range(B):
            o = int(off[b])
            for c in range(C):
                k = int(n[b, c])
                k4 = (k + 3) & ~3
                out[b, c, :k] = pk[o:o + k]
                out[b, C + c, :k] = 1.0
                out[b, 2 * C + c, :k] = pk[o + k4:o + k4 + k]
                o += 2 * k4
        return out


class PackedStaging:
    """Device-side landing buffers for one chunk of packed encounters."""

    def __init__(self, C, max_encounters, max_floats, device=None):
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.C, self.cap_b, self.cap_f = int(C), int(max_encounters), int(max_floats)
        self.packed = torch.empty(max(self.cap_f, 4), dtype=torch.float32, device=dev)
        self.n_obs = torch.empty((self.cap_b, self.C), dtype=torch.int32, device=dev)
        self.enc_off = torch.empty(self.cap_b + 1, dtype=torch.int64, device=dev)
        self.device = dev

    @classmethod
    def for_chunks(cls, pe, chunk, device=None):
        return cls(pe.C, min(chunk, max(pe.B, 1)), pe.max_chunk_floats(chunk), device)

    def upload(self, pe, b0=0, b1=None, out=None, stream=None, dev_planes=None):
        """H2D of encounters [b0, b1) of `pe` + expansion to dense planes.  Returns x (b1-b0, dev_planes, T) on the
        device (dev_planes = 3C by default), usable wherever the reference's x is (SingleChannelInterp, RBF)."""
        b1 = pe.B if b1 is None else b1
        n = b1 - b0
        if n < 0 or b0 < 0 or b1 > pe.B:
            raise ValueError(f"bad encounter range [{b0}, {b1}) of {pe.B}")
        if n > self.cap_b or pe.floats(b0, b1) > self.cap_f or pe.C != self.C:
            raise ValueError(f"chunk of {n} encounters / {pe.floats(b0, b1)} floats exceeds the staging capacity "
                             f"({self.cap_b} / {self.cap_f})")
        P = 3 * pe.C if dev_planes is None else int(dev_planes)
        if out is None:
            out = torch.empty((n, P, pe.T), dtype=torch.float32, device=self.device)
        if tuple(out.shape) != (n, P, pe.T) or not out.is_contiguous() or out.device != self.device \
                or out.dtype != torch.float32:
            raise ValueError(f"out must be a dense float32 tensor of shape {(n, P, pe.T)} on {self.device}")
        st = (stream or torch.cuda.current_stream(self.device)).cuda_stream
        L = _lib.lib()
        with torch.cuda.device(self.device):
            _lib.check(L.dic_upload_encounters_packed(
                pe.packed.data_ptr(), pe.n_obs.data_ptr() + 4 * b0 * pe.C, pe.enc_off.data_ptr() + 8 * b0,
                _lib.ptr(self.packed), _lib.ptr(self.n_obs), _lib.ptr(self.enc_off), n, pe.C, st),
                "dic_upload_encounters_packed")
            _lib.check(L.dic_expand_encounters(_lib.ptr(self.packed), _lib.ptr(self.n_obs), _lib.ptr(self.enc_off),
                                               _lib.ptr(out), n, pe.C, pe.T, P, st), "dic_expand_encounters")
        return out
