"""The hot-path operators chained inside the network they serve (pretrain_interp.py:138-142, 169-175 and
clustering_interp.py:186-207): SCI -> CCI -> BiLSTM encoder -> BiLSTM decoder -> compress_fc -> RBF read-out,
masked reconstruction loss + 10 x the DEC KL term, ONE backward pass.

The B200 arm is what a user of the reference runs after swapping the modules in (torch's LSTM between the
custom autograd functions, float32); the truth is the same network assembled from the oracle's torch port of the
reference operators on the CPU in float64 with the same weights.  Checks the plumbing no per-operator test sees:
gradients flowing RBF -> compress_fc -> LSTMs -> CCI -> SCI through the custom backward kernels, the (B, R, 3C)
permuted view handed to the LSTM, and the DEC branch hanging off the encoder state.
"""
import copy

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

C, T, R, H, NH, K = 6, 32, 24, 24.0, 16, 3


class _Net(nn.Module):
    """Mirror of pretrain_interp.Net.forward (:138-142) around whatever operator modules it is given."""

    def __init__(self, sci, cci, rbf, assign, nh=NH, lstm=None):
        super().__init__()
        self.sci, self.cci, self.rbf, self.assign = sci, cci, rbf, assign
        lstm = lstm or (lambda i, h: nn.LSTM(i, h, bidirectional=True))
        self.encoder = lstm(3 * C, nh)
        self.decoder = lstm(2 * nh, nh)

    def forward(self, x):
        u = self.cci(self.sci(x)).permute(1, 0, 2)                 # (R, B, 3C), as the LSTM wants it
        enc, (h, c) = self.encoder(u)
        y, _ = self.decoder(F.relu(enc), (h, c))
        rec = self.rbf(y.permute(1, 2, 0), x)                      # (B, C, T)
        hidden = torch.cat([h[0], h[1]], dim=-1)                   # (B, 2 NH)
        return hidden, rec


def _loss(x, hidden, rec, q_fn, p_fn):
    ob, m = x[:, :C], x[:, C:2 * C]
    mse = ((rec * m - ob * m) ** 2).sum() / (m == 1.0).sum()       # pretrain_interp.py:169-175
    q = q_fn(hidden)
    p = p_fn(q).detach()                                           # clustering_interp.py:186
    return mse + 10.0 * F.kl_div(q.log(), p, reduction="batchmean"), q


def test_chained_network_step_matches_float64_reference_operators(monkeypatch):
    _chained_step(monkeypatch, NH, None)


def test_chained_network_with_the_b200_bilstm_matches_float64_reference(monkeypatch):
    """The same step with the encoder / decoder LSTMs on the persistent tcgen05 kernel (lstm.BiLSTMB200, hidden 128 as in
    pretrain_interp.py:95-112): every operator between the observations and the loss is now this package's."""
    from deep_interpolation_clustering_b200.lstm import BiLSTMB200
    _chained_step(monkeypatch, 128, lambda i, h: BiLSTMB200(i, h))


def _chained_step(monkeypatch, NH, b200_lstm):
    import deep_interpolation_clustering_b200 as dic
    from deep_interpolation_clustering_b200 import synth
    from oracle import ref_port

    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    # plain float32 arithmetic in the library layers: no TF32, and ATen's LSTM instead of cuDNN's fused one, whose
    # approximate gate activations alone put 5e-4 ... 1e-3 on every LSTM gradient (measured) and would hide a
    # plumbing error of that size in the operators under test
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    monkeypatch.setattr(torch.backends.cuda.matmul, "allow_tf32", False)
    monkeypatch.setattr(torch.backends.cudnn, "enabled", False)
    x_np = synth.make_encounters(48, C, T, H, seed=3)
    rng = np.random.RandomState(1)

    sci = dic.SingleChannelInterp(R, H, C, T, dev)
    cci = dic.CrossChannelInterp(C, T, dev)
    rbf = dic.RBF(H, R, 2 * NH, C, 0.0, dic.basis_func_dict()["gaussian"], dev)
    ca = dic.ClusterAssignment(K, 2 * NH, 1.0)
    with torch.no_grad():
        cci.kernel.add_(torch.from_numpy(0.1 * rng.standard_normal((C, C)).astype(np.float32)).to(cci.kernel.device))
        ca.cluster_centers.copy_(torch.from_numpy(0.3 * rng.standard_normal((K, 2 * NH)).astype(np.float32)))
    ours = _Net(sci, cci, rbf, ca, NH, b200_lstm)

    # the truth: oracle operators, same weights, float64 on the CPU
    o_sci = ref_port.SingleChannelInterp(R, H, C)
    o_cci = ref_port.CrossChannelInterp(C)
    o_rbf = ref_port.RBFReadout(R, H, C, compress=copy.deepcopy(rbf.compress_fc.module.model).cpu())
    truth = _Net(o_sci, o_cci, o_rbf, None, NH)
    with torch.no_grad():
        o_sci.kernel.copy_(sci.kernel)
        o_cci.kernel.copy_(cci.kernel)
        o_rbf.kernel.copy_(rbf.kernel)
    truth.encoder.load_state_dict(ours.encoder.state_dict())
    truth.decoder.load_state_dict(ours.decoder.state_dict())
    truth = truth.double()
    mu64 = ca.cluster_centers.detach().cpu().double().clone().requires_grad_(True)

    ours = ours.to(dev).train()
    truth.train()
    x = torch.from_numpy(x_np).to(dev)
    hidden, rec = ours(x)
    loss, q = _loss(x, hidden, rec, ours.assign, dic.target_distribution)
    loss.backward()

    x64 = torch.from_numpy(x_np).double()
    hidden64, rec64 = truth(x64)
    loss64, q64 = _loss(x64, hidden64, rec64, lambda z: ref_port.soft_assign(z, mu64), ref_port.target_distribution)
    loss64.backward()

    assert abs(float(loss) - float(loss64)) <= 2e-5 * abs(float(loss64)), (float(loss), float(loss64))
    assert torch.allclose(q.detach().cpu().double(), q64.detach(), rtol=1e-4, atol=1e-6)
    assert torch.allclose(rec.detach().cpu().double(), rec64.detach(), rtol=1e-4, atol=2e-5)

    report, bad = [], []

    def close(name, got, want, rel):
        got, want = got.detach().cpu().double(), want.detach().double()
        err = float((got - want).abs().max())
        scale = float(want.abs().max())
        if scale < 1e-10:          # e.g. the bias in front of BatchNorm: its gradient vanishes analytically
            ok = err < 1e-6
        else:
            ok = err <= rel * scale
        report.append(f"{name}: max err {err:.3e}, scale {scale:.3e}, ratio {err / max(scale, 1e-30):.2e} (limit {rel:g})")
        if not ok:
            bad.append(report[-1])

    # the path's own parameters (a float32 LSTM sits between them and the loss).  Measured: 5e-7 ... 3e-6 of the
    # largest entry - for d sci.kernel, a heavily cancelling sum, the reference's own operators run in float32 on
    # the CPU are 6.5e-4 off the float64 value on this very network.
    close("d sci.kernel", sci.kernel.grad, o_sci.kernel.grad, 5e-5)
    close("d cci.kernel", cci.kernel.grad, o_cci.kernel.grad, 5e-5)
    close("d rbf.kernel", rbf.kernel.grad, o_rbf.kernel.grad, 5e-5)
    close("d cluster_centers", ca.cluster_centers.grad, mu64.grad, 5e-5)
    # the library layers in between see the right upstream gradients
    for (name, p), (_, p64) in zip(rbf.compress_fc.module.model.named_parameters(),
                                   o_rbf.compress.named_parameters()):
        close("d compress_fc." + name, p.grad, p64.grad, 5e-5)
    for part in ("encoder", "decoder"):
        for (name, p), (_, p64) in zip(getattr(ours, part).named_parameters(), getattr(truth, part).named_parameters()):
            close(f"d {part}.{name}", p.grad, p64.grad, 5e-5)
    print("\n".join(report))
    assert not bad, "\n".join(bad)
