"""Ragged (packed) encounter format: host packer on the CPU, upload + expansion and operator parity on the GPU.

Reference context: the pipeline's rows are left-packed (p0_data_process.py:44-67) and the trainer ships them as
dense planes (pretrain_trainer.py:132-136); the packed path must deliver bit-identical operator results.
"""
import numpy as np
import pytest
import torch

import deep_interpolation_clustering_b200 as dic
from deep_interpolation_clustering_b200 import PackedEncounters, PackedStaging, synth
from deep_interpolation_clustering_b200 import functional as F_


@pytest.mark.parametrize("B,C,T", [(37, 6, 64), (5, 3, 30), (9, 6, 256), (4, 1, 7), (3, 16, 12)])
def test_host_packer_round_trip(B, C, T):
    x = synth.make_encounters(B, C, T, 24.0, seed=B + T)
    pe = PackedEncounters.from_dense(x, pin=False)
    assert pe.B == B and pe.C == C and pe.T == T and pe.all_sorted
    n = x[:, C:2 * C].sum(-1).astype(np.int64)
    assert np.array_equal(pe.n_obs.numpy(), n)
    off = pe.enc_off.numpy()
    assert off[0] == 0 and np.all(off % 4 == 0)
    assert np.array_equal(np.diff(off), (2 * ((n + 3) // 4 * 4)).sum(1))
    assert np.array_equal(pe.to_dense(), x[:, :3 * C])
    # pad slots: value 0, time 3e18 (weight exactly 0 in every Gaussian sum)
    pk = pe.packed.numpy()
    k, k4 = int(n[0, 0]), int((n[0, 0] + 3) // 4 * 4)
    assert np.all(pk[k:k4] == 0.0) and np.all(pk[k4 + k:2 * k4] == np.float32(3.0e18))
    # three-plane input packs to the same thing
    pe3 = PackedEncounters.from_dense(np.ascontiguousarray(x[:, :3 * C]), pin=False) if C % 4 else None
    if pe3 is not None:
        assert np.array_equal(pe3.packed.numpy(), pk)
    assert pe.nbytes() == 4 * int(off[-1]) + 4 * B * C + 8 * (B + 1)


def test_host_packer_edge_cases():
    C, T = 2, 8
    x = np.zeros((3, 4 * C, T), np.float32)
    x[0, C:2 * C] = 1.0                      # full rows
    x[0, 2 * C:3 * C] = np.arange(T)
    x[1, C, :3] = 1.0                        # one vital with 3 observations, the other all-masked
    x[1, 2 * C, :3] = [5.0, 2.0, 9.0]        # unsorted times are kept in slot order (RBF indexes by slot)
    pe = PackedEncounters.from_dense(x, pin=False)
    assert pe.n_obs.numpy().tolist() == [[8, 8], [3, 0], [0, 0]]
    assert not pe.all_sorted
    assert np.array_equal(pe.to_dense(), x[:, :3 * C])
    assert pe.floats(2, 3) == 0
    empty = PackedEncounters.from_dense(np.zeros((0, 4 * C, T), np.float32), pin=False)
    assert empty.B == 0 and empty.floats() == 0


def test_host_packer_rejects_general_masks():
    xa = synth.make_adversarial_encounters(4, 6, 30)
    with pytest.raises(ValueError, match="left-packed"):
        PackedEncounters.from_dense(xa, pin=False)
    xw = synth.make_encounters(2, 6, 16)
    xw[0, 6, 0] = 0.5                        # fractional weight
    with pytest.raises(ValueError, match="left-packed"):
        PackedEncounters.from_dense(xw, pin=False)


@pytest.mark.gpu
@pytest.mark.parametrize("B,C,T,R", [(64, 6, 256, 96), (33, 6, 64, 48), (17, 5, 30, 40), (8, 6, 1024, 192)])
def test_packed_upload_is_bit_identical_to_dense(B, C, T, R):
    dev = torch.device("cuda:0")
    xn = synth.make_encounters(B, C, T, 24.0, seed=7)
    xh = torch.from_numpy(xn).pin_memory()
    dense = F_.upload_encounters(xh, device=dev)
    pe = PackedEncounters.from_dense(xh)
    st = PackedStaging.for_chunks(pe, B, dev)
    x = st.upload(pe)
    assert torch.equal(x, dense)
    # a chunked upload of a sub-range lands the same rows
    b0, b1 = B // 3, B - 2
    st2 = PackedStaging(C, b1 - b0, pe.floats(b0, b1), dev)
    assert torch.equal(st2.upload(pe, b0, b1), dense[b0:b1])
    # expansion into a 4C-plane buffer leaves the fourth plane alone
    buf = torch.full((B, 4 * C, T), -7.0, device=dev)
    st.upload(pe, out=buf, dev_planes=4 * C)
    assert torch.equal(buf[:, :3 * C], dense) and bool((buf[:, 3 * C:] == -7.0).all())

    p = synth.make_interp_params(C, seed=1)
    sci = dic.SingleChannelInterp(R, 24.0, C, T, dev)
    cci = dic.CrossChannelInterp(C, T, dev)
    rbf = dic.RBF(24.0, R, C, C, 0.0, dic.basis_func_dict()["gaussian"], dev)
    rbf.compress_fc = torch.nn.Identity()
    sci.kernel.data = torch.tensor(p["sci_kernel"], device=dev)
    cci.kernel.data = torch.tensor(p["cci_kernel"], device=dev)
    rbf.kernel.data = torch.tensor(p["rbf_kernel"], device=dev)
    v = torch.randn((B, C, R), device=dev, generator=torch.Generator(device=dev).manual_seed(3))
    res = []
    for xin in (torch.from_numpy(xn).to(dev), x):
        for m in (sci, cci, rbf):
            m.zero_grad()
        vv = v.clone().requires_grad_(True)
        out = cci(sci(xin))
        rec = rbf(vv, xin)
        (out.square().sum() + rec.square().sum()).backward()
        res.append([out.detach().clone(), rec.detach().clone(), vv.grad.clone(), sci.kernel.grad.clone(),
                    cci.kernel.grad.clone(), rbf.kernel.grad.clone()])
    for a, b in zip(*res):
        assert torch.equal(a, b)
