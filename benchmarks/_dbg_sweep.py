import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import deep_interpolation_clustering_b200 as dic
from deep_interpolation_clustering_b200 import synth
from oracle import interp_oracle as O
dev = torch.device("cuda:0")
def run(B, C, T, R, seed_off=0):
    H = 24.0
    xn = synth.make_encounters(B, C, T, H, seed=B * 1000 + T + seed_off)
    rng = np.random.RandomState(C * 100 + R + seed_off)
    ks = rng.uniform(size=C).astype(np.float32)
    kc = (np.eye(C) + 0.1 * rng.normal(size=(C, C))).astype(np.float32)
    gc = rng.normal(size=(B, R, 3 * C)).astype(np.float32)
    rt = O.linspace_grid(H, R)
    x64 = xn.astype(np.float64)
    s64 = O.sci_forward(x64, ks.astype(np.float64), rt, C)
    du64, _ = O.cci_backward(s64, kc.astype(np.float64), C, gc.astype(np.float64))
    dks64 = O.sci_backward(x64, ks.astype(np.float64), rt, C, du64)
    # float32 oracle as a yardstick of what float32 arithmetic gives on this problem
    s32 = O.sci_forward(xn, ks, rt, C)
    du32, _ = O.cci_backward(s32, kc, C, gc)
    dks32 = O.sci_backward(xn, ks, rt, C, du32)
    sci = dic.SingleChannelInterp(R, H, C, T, dev); cci = dic.CrossChannelInterp(C, T, dev)
    sci.kernel.data, cci.kernel.data = torch.tensor(ks, device=dev), torch.tensor(kc, device=dev)
    c = cci(sci(torch.tensor(xn, device=dev)))
    (c * torch.tensor(gc, device=dev)).sum().backward()
    got = sci.kernel.grad.cpu().numpy().astype(np.float64)
    rms = np.sqrt((dks64 ** 2).mean())
    print(f"B={B} C={C} T={T} R={R}: gpu err {np.abs(got - dks64).max():.2e}  f32-oracle err {np.abs(dks32 - dks64).max():.2e}  rms {rms:.2e}")
for args in [(3, 10, 20, 16), (30, 10, 20, 16), (3, 9, 20, 16), (3, 6, 20, 16), (3, 8, 20, 16), (3, 10, 20, 16, 5), (3, 10, 64, 48), (300, 10, 20, 16)]:
    run(*args)
