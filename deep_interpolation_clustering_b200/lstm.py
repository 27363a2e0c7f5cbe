"""Encoder / decoder BiLSTM mirrors.

The reference wraps ``nn.LSTM(input, 128, num_layers=1, bidirectional=True)`` in ``EncoderRNN`` / ``DecoderRNN``
(pretrain_interp.py:14-41; the same classes in clustering_interp.py) and runs it time-major on the CCI output
(R, B, 18) and on the encoder output (R, B, 256).  ``BiLSTMB200`` keeps nn.LSTM's parameter names and shapes
(``weight_ih_l0 (512, I)``, ``weight_hh_l0 (512, 128)``, ``bias_ih_l0``, ``bias_hh_l0`` and the ``_reverse`` set), so
``encoder.lstm.*`` / ``decoder.lstm.*`` checkpoints load unchanged, and computes

  * the input projection of all steps and both directions in ONE tcgen05 kernel with the recurrence's arithmetic
    (``dic_lstm_project``: fp16 hi + lo operands split on the fly, float32 accumulation in TMEM; the float32 SIMT
    library GEMM it replaces cost 2-4x the whole recurrence, and the library's TF32 tensor-core GEMMs are 3-8x less
    accurate - their accumulation is not round-to-nearest - i.e. outside the 1e-5 parity, profiles/r02_lstm_notes.txt),
  * the recurrence of all R steps in ONE persistent sm_100a kernel (``dic_lstm_fwd``: 4-CTA clusters, W_hh resident
    in shared memory, h exchanged over distributed shared memory, tcgen05 MMAs on split-fp16 operands),
  * the backward recurrence with one fused gate-gradient kernel per step (``dic_lstm_bwd_step``) and library GEMMs
    for d h_(t-1) = d a_t W_hh and for the weight / input gradients.

``dropin.patch_lstm(module)`` rebinds ``EncoderRNN`` / ``DecoderRNN`` of an imported ``pretrain_interp`` /
``clustering_interp`` to the mirrors below, so the unmodified ``Net`` builds on them.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib

H = 128


def _perm_index(device):
    """Column order of ``pre`` expected by dic_lstm_fwd: [direction][slice q][half][gate][16 units] -> row of the
    stacked [weight_ih_l0 ; weight_ih_l0_reverse] (gate-major rows i|f|g|o as in torch)."""
    d, q, h, g, u = torch.meshgrid(torch.arange(2), torch.arange(4), torch.arange(2), torch.arange(4), torch.arange(16),
                                   indexing="ij")
    return (d * 4 * H + g * H + 32 * q + 16 * h + u).reshape(-1).to(device)


class _BiLSTM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, h0, c0, w_ih, w_hh, b_ih, b_hh, w_ih_r, w_hh_r, b_ih_r, b_hh_r):
        R, B, I = x.shape
        dev = x.device
        L = _lib.lib()
        need_grad = any(ctx.needs_input_grad)
        with torch.cuda.device(dev):
            st = _lib.current_stream(dev)
            perm = _perm_index(dev)
            wp = torch.cat([w_ih, w_ih_r], 0)[perm]                                   # (1024, I)
            bp = torch.cat([b_ih + b_hh, b_ih_r + b_hh_r], 0)[perm]
            # the input projection of all steps and both directions: split-fp16 tcgen05 GEMM (dic_lstm_project)
            wp, bp, x2 = wp.contiguous(), bp.contiguous(), x.reshape(R * B, I)
            pw = torch.empty(int(L.dic_lstm_project_packed_bytes(I)), dtype=torch.uint8, device=dev)
            _lib.check(L.dic_lstm_pack_wih(_lib.ptr(wp), _lib.ptr(pw), I, st), "dic_lstm_pack_wih")
            pre = torch.empty((R * B, 8 * H), dtype=torch.float32, device=dev)
            _lib.check(L.dic_lstm_project(_lib.ptr(x2), I, _lib.ptr(pw), _lib.ptr(bp), _lib.ptr(pre), R * B, I, 0, st),
                       "dic_lstm_project")
            packed = torch.empty(int(L.dic_lstm_packed_bytes()), dtype=torch.uint8, device=dev)
            w_hh, w_hh_r = w_hh.contiguous(), w_hh_r.contiguous()
            _lib.check(L.dic_lstm_pack_whh(_lib.ptr(w_hh), _lib.ptr(w_hh_r), _lib.ptr(packed), st), "dic_lstm_pack_whh")
            out = torch.empty((R, B, 2 * H), dtype=torch.float32, device=dev)
            hn = torch.empty((2, B, H), dtype=torch.float32, device=dev)
            cn = torch.empty((2, B, H), dtype=torch.float32, device=dev)
            save = torch.empty((2, R, B, 5, H), dtype=torch.float32, device=dev) if need_grad else None
            _lib.check(L.dic_lstm_fwd(_lib.ptr(pre), _lib.ptr(packed), _lib.ptr(h0), _lib.ptr(c0), _lib.ptr(out),
                                      _lib.ptr(hn), _lib.ptr(cn), _lib.ptr(save), R, B, H, st), "dic_lstm_fwd")
        if need_grad:
            ctx.save_for_backward(x, h0, c0, w_ih, w_hh, w_ih_r, w_hh_r, out, save)
        return out, hn, cn

    @staticmethod
    def backward(ctx, g_out, g_hn, g_cn):
        x, h0, c0, w_ih, w_hh, w_ih_r, w_hh_r, out, save = ctx.saved_tensors
        R, B, I = x.shape
        dev = x.device
        L = _lib.lib()
        g_out = None if g_out is None else g_out.contiguous()
        grads_w = []
        dx = torch.zeros((R * B, I), dtype=torch.float32, device=dev)
        dh0 = torch.empty((2, B, H), dtype=torch.float32, device=dev)
        dc0 = torch.empty((2, B, H), dtype=torch.float32, device=dev)
        x2 = x.reshape(R * B, I)
        with torch.cuda.device(dev):
            st = _lib.current_stream(dev)
            for d, (wi, wh) in enumerate(((w_ih, w_hh), (w_ih_r, w_hh_r))):
                dh = (g_hn[d].clone() if g_hn is not None else torch.zeros((B, H), device=dev)).contiguous()
                dc = (g_cn[d].clone() if g_cn is not None else torch.zeros((B, H), device=dev)).contiguous()
                da = torch.empty((R, B, 4 * H), dtype=torch.float32, device=dev)
                order = range(R - 1, -1, -1) if d == 0 else range(R)          # reverse of the processing order
                for t in order:
                    tp = t - 1 if d == 0 else t + 1                            # the step processed just before t
                    if 0 <= tp < R:
                        c_prev, c_stride = save[d, tp, :, 4], 5 * H
                    elif c0 is not None:
                        c_prev, c_stride = c0[d], H
                    else:
                        c_prev, c_stride = None, 0
                    gh = g_out[t, :, d * H:] if g_out is not None else None
                    _lib.check(L.dic_lstm_bwd_step(_lib.ptr(save[d, t]), _lib.ptr(c_prev), c_stride, _lib.ptr(gh), 2 * H,
                                                   _lib.ptr(dh), _lib.ptr(dc), _lib.ptr(da[t]), B, H, st),
                               "dic_lstm_bwd_step")
                    torch.mm(da[t], wh, out=dh)                                # d h_(t-1) = d a_t W_hh (library GEMM)
                dh0[d], dc0[d] = dh, dc
                da2 = da.reshape(R * B, 4 * H)
                # h_(t-1) of every step: the layer output shifted by one step in processing order (+ h0 at the first)
                if d == 0:
                    dwh = da[1:].reshape(-1, 4 * H).t() @ out[:-1, :, :H].reshape(-1, H) if R > 1 else 0
                    first = 0
                else:
                    dwh = da[:-1].reshape(-1, 4 * H).t() @ out[1:, :, H:].reshape(-1, H) if R > 1 else 0
                    first = R - 1
                if h0 is not None:
                    dwh = dwh + da[first].t() @ h0[d]
                if not isinstance(dwh, torch.Tensor):
                    dwh = torch.zeros_like(wh)
                db = da2.sum(0)
                grads_w.append((da2.t() @ x2, dwh, db, db))
                dx.addmm_(da2, wi)
                del da, da2
        (gi, gh_, gbi, gbh), (gir, ghr, gbir, gbhr) = grads_w
        need = ctx.needs_input_grad
        return (dx.view(R, B, I) if need[0] else None, dh0 if need[1] else None, dc0 if need[2] else None,
                gi, gh_, gbi, gbh, gir, ghr, gbir, gbhr)


class BiLSTMB200(nn.Module):
    """``nn.LSTM(input_size, 128, num_layers=1, bidirectional=True)`` on the B200 kernels: same parameters, same
    ``forward(x (R, B, I)[, (h0, c0) (2, B, 128)]) -> (output (R, B, 256), (h_n, c_n))``, time-major like the reference."""

    def __init__(self, input_size, hidden_size=H, num_layers=1, bias=True, batch_first=False, dropout=0.0,
                 bidirectional=True):
        super().__init__()
        if hidden_size != H or num_layers != 1 or not bidirectional or not bias or batch_first:
            raise ValueError("BiLSTMB200 covers the reference's configuration: hidden_size=128, num_layers=1, "
                             "bidirectional=True, bias=True, time-major input (pretrain_interp.py:95-112)")
        self.input_size, self.hidden_size, self.num_layers, self.bidirectional = input_size, hidden_size, 1, True
        self.dropout = dropout                      # nn.LSTM applies dropout between layers only: none with one layer
        for suffix in ("", "_reverse"):
            self.register_parameter("weight_ih_l0" + suffix, nn.Parameter(torch.empty(4 * H, input_size)))
            self.register_parameter("weight_hh_l0" + suffix, nn.Parameter(torch.empty(4 * H, H)))
            self.register_parameter("bias_ih_l0" + suffix, nn.Parameter(torch.empty(4 * H)))
            self.register_parameter("bias_hh_l0" + suffix, nn.Parameter(torch.empty(4 * H)))
        self.reset_parameters()

    def reset_parameters(self):
        k = 1.0 / math.sqrt(H)                      # nn.LSTM.reset_parameters
        for p in self.parameters():
            nn.init.uniform_(p, -k, k)

    def forward(self, x, hx=None):
        if not x.is_cuda:
            raise RuntimeError("BiLSTMB200 runs on a CUDA device only (sm_100a kernels; no CPU fallback)")
        if x.dim() != 3 or x.shape[2] != self.input_size:
            raise RuntimeError(f"input must be (seq, batch, {self.input_size}); got {tuple(x.shape)}")
        x = x.contiguous().float()
        h0 = c0 = None
        if hx is not None:
            h0, c0 = hx
            if tuple(h0.shape) != (2, x.shape[1], H) or tuple(c0.shape) != (2, x.shape[1], H):
                raise RuntimeError(f"Expected hidden size (2, {x.shape[1]}, {H}), got {tuple(h0.shape)}")
            h0, c0 = h0.contiguous().float(), c0.contiguous().float()
        out, hn, cn = _BiLSTM.apply(x, h0, c0, self.weight_ih_l0, self.weight_hh_l0, self.bias_ih_l0, self.bias_hh_l0,
                                    self.weight_ih_l0_reverse, self.weight_hh_l0_reverse, self.bias_ih_l0_reverse,
                                    self.bias_hh_l0_reverse)
        return out, (hn, cn)


class EncoderRNN(nn.Module):
    """pretrain_interp.py:14-27 with the LSTM on the B200 kernels (same constructor, same ``lstm.*`` state-dict keys)."""

    def __init__(self, input_size, hidden_size, num_layers, dropout, bidirectional, device):
        super().__init__()
        self.device = device
        self.hidden_size = hidden_size
        self.num_layers = num_layers
        self.num_directions = 2 if bidirectional else 1
        self.lstm = BiLSTMB200(input_size, hidden_size, num_layers=num_layers, dropout=dropout, bidirectional=bidirectional)

    def forward(self, x):
        output, (hidden, cell_state) = self.lstm(x)
        return output, hidden, cell_state


class DecoderRNN(nn.Module):
    """pretrain_interp.py:29-41."""

    def __init__(self, input_size, hidden_size, num_layers, dropout, bidirectional, device):
        super().__init__()
        self.device = device
        self.hidden_size = hidden_size
        self.lstm = BiLSTMB200(input_size, hidden_size, num_layers=num_layers, dropout=dropout, bidirectional=bidirectional)

    def forward(self, x, hidden, context):
        x = F.relu(x)
        x, (hidden, cell_state) = self.lstm(x, (hidden, context))
        return x, (hidden, cell_state)
