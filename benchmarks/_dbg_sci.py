import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deep_interpolation_clustering_b200 import synth, functional as F_
from oracle import interp_oracle
B, C, T, R, H = 4096, 6, 256, 96, 24.0
xn = synth.make_encounters(B, C, T, H, seed=0)
p = synth.make_interp_params(C, seed=1)
rt = interp_oracle.linspace_grid(H, R)
dev = torch.device("cuda:0")
u = F_.sci(torch.tensor(xn, device=dev), torch.tensor(p["sci_kernel"], device=dev), torch.tensor(rt.astype(np.float32), device=dev)).cpu().numpy()
worst = (0, None)
for i in range(0, B, 256):
    s64 = interp_oracle.sci_forward(xn[i:i + 256].astype(np.float64), p["sci_kernel"].astype(np.float64), rt, C)   # (b, R, 3C)
    s64 = np.transpose(s64, (0, 2, 1))      # (b, 3C, R)
    err = np.abs(u[i:i + 256] - s64)
    for g in range(3):
        e = err[:, g * C:(g + 1) * C]
        k = np.unravel_index(np.argmax(e), e.shape)
        if g == 2 and e[k] > worst[0]:
            worst = (e[k], (i + k[0], k[1], k[2], u[i + k[0], 2 * C + k[1], k[2]], s64[k[0], 2 * C + k[1], k[2]]))
    print(i, [float(err[:, g * C:(g + 1) * C].max()) for g in range(3)], flush=True) if i % 1024 == 0 else None
print("worst high-pass", worst)
e, (b, c, r, got, want) = worst
m = xn[b, C + c] > 0
d = xn[b, 2 * C + c][m]; xv = xn[b, c][m]
print("n obs", m.sum(), "r =", rt[r], "alpha", np.log1p(np.exp(p["sci_kernel"][c])))
near = np.argsort(np.abs(d - rt[r]))[:8]
for j in sorted(near): print("  d", d[j], "delta", d[j] - rt[r], "x", xv[j])
