"""GPU parity of the BiLSTM mirrors against torch.nn.LSTM evaluated in float64 on the host with the same
weights: outputs, final states and EVERY gradient (input, initial states, all eight parameter tensors) to rtol 1e-5
(+ 1e-5 of the tensor's RMS as the absolute floor, like the interpolation tests)."""
import numpy as np
import pytest
import torch

from gpu_util import record, scale_atol

pytestmark = pytest.mark.gpu


def _reference(lstm_b200, x, hx, g_out, g_hn, g_cn):
    I = lstm_b200.input_size
    ref = torch.nn.LSTM(I, 128, num_layers=1, bidirectional=True).double()
    ref.load_state_dict({k: v.detach().cpu().double() for k, v in lstm_b200.state_dict().items()})
    x64 = x.detach().cpu().double().requires_grad_(True)
    hx64 = None
    if hx is not None:
        hx64 = tuple(t.detach().cpu().double().requires_grad_(True) for t in hx)
    out, (hn, cn) = ref(x64, hx64)
    ((out * g_out.cpu().double()).sum() + (hn * g_hn.cpu().double()).sum() + (cn * g_cn.cpu().double()).sum()).backward()
    return ref, x64, hx64, out, hn, cn


def _check(name, got, want, rtol=1e-5):
    want = want.detach().numpy()
    if not np.any(want):            # e.g. d weight_hh of a one-step sequence without an initial state: exactly zero
        assert not np.any(got.detach().cpu().numpy()), f"{name}: expected exact zeros"
        return None
    return record(name, got.detach().cpu().numpy(), want, rtol, scale_atol(want, rtol))


@pytest.mark.parametrize("R,B,I,with_state", [(5, 7, 18, False), (12, 130, 18, False), (9, 300, 256, True),
                                               (96, 257, 18, False), (48, 64, 256, True), (1, 1, 18, False),
                                               (2, 129, 256, True), (3, 128, 100, True)])
def test_bilstm_matches_torch_float64(R, B, I, with_state):
    from deep_interpolation_clustering_b200.lstm import BiLSTMB200
    torch.manual_seed(R * 1000 + B)
    dev = torch.device("cuda:0")
    m = BiLSTMB200(I).to(dev)
    x = (torch.randn(R, B, I, device=dev) * (1.0 if I == 18 else 0.5)).requires_grad_(True)
    hx = None
    if with_state:
        hx = (torch.tanh(torch.randn(2, B, 128, device=dev)).requires_grad_(True),
              torch.randn(2, B, 128, device=dev).requires_grad_(True))
    g_out, g_hn, g_cn = torch.randn(R, B, 256, device=dev), torch.randn(2, B, 128, device=dev), torch.randn(2, B, 128, device=dev)
    out, (hn, cn) = m(x, hx)
    assert out.shape == (R, B, 256) and hn.shape == (2, B, 128) and cn.shape == (2, B, 128)
    ((out * g_out).sum() + (hn * g_hn).sum() + (cn * g_cn).sum()).backward()
    ref, x64, hx64, out64, hn64, cn64 = _reference(m, x, hx, g_out, g_hn, g_cn)
    tag = f"lstm_R{R}_B{B}_I{I}"
    _check(tag + "/out", out, out64)
    _check(tag + "/hn", hn, hn64)
    _check(tag + "/cn", cn, cn64)
    _check(tag + "/dx", x.grad, x64.grad)
    if with_state:
        _check(tag + "/dh0", hx[0].grad, hx64[0].grad)
        _check(tag + "/dc0", hx[1].grad, hx64[1].grad)
    for k, p in m.named_parameters():
        _check(f"{tag}/d_{k}", p.grad, dict(ref.named_parameters())[k].grad)


def test_bilstm_inference_needs_no_saved_state_and_matches_training_forward():
    from deep_interpolation_clustering_b200.lstm import BiLSTMB200
    torch.manual_seed(3)
    dev = torch.device("cuda:0")
    m = BiLSTMB200(18).to(dev)
    x = torch.randn(20, 200, 18, device=dev)
    with torch.no_grad():
        a, (ha, ca) = m(x)
    b, (hb, cb) = m(x.clone().requires_grad_(True))
    assert torch.equal(a, b) and torch.equal(ha, hb) and torch.equal(ca, cb)
    assert torch.equal(a[-1, :, :128], ha[0]) and torch.equal(a[0, :, 128:], ha[1])      # nn.LSTM's h_n convention


def test_encoder_decoder_mirrors_keep_the_reference_state_dict_keys():
    from deep_interpolation_clustering_b200.lstm import DecoderRNN, EncoderRNN
    dev = torch.device("cuda:0")
    enc = EncoderRNN(18, 128, num_layers=1, dropout=0, bidirectional=True, device=dev).to(dev)
    dec = DecoderRNN(input_size=256, hidden_size=128, num_layers=1, dropout=0, bidirectional=True, device=dev).to(dev)
    want = sorted(torch.nn.LSTM(18, 128, bidirectional=True).state_dict())
    assert sorted(k[len("lstm."):] for k in enc.state_dict()) == want
    x = torch.randn(10, 33, 18, device=dev)
    o, h, c = enc(x)
    y, (h2, c2) = dec(o, h, c)
    assert y.shape == (10, 33, 256) and h2.shape == (2, 33, 128)
    with pytest.raises(ValueError):
        EncoderRNN(18, 64, num_layers=1, dropout=0, bidirectional=True, device=dev)
