"""Per-kernel timings of the interpolation path at the c2 shape (B encounters, default 262144): CUDA events, best of n.
    python benchmarks/_interp_kernels.py [B] [T] [R]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_interpolation_clustering_b200 import functional as F_, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
T = int(sys.argv[2]) if len(sys.argv) > 2 else 256
R = int(sys.argv[3]) if len(sys.argv) > 3 else 96
C, H = 6, 24.0
dev = torch.device("cuda:0")
x = synth.make_encounters_device(B, C, T, H, 5.0, 0, dev)
p = synth.make_interp_params(C, seed=1)
ks = torch.tensor(p["sci_kernel"], device=dev, requires_grad=True)
kc = torch.tensor(p["cci_kernel"], device=dev, requires_grad=True)
kr = torch.tensor(p["rbf_kernel"], device=dev, requires_grad=True)
rt = torch.linspace(0, H, R, device=dev)
v = torch.randn(B, C, R, device=dev, requires_grad=True)
g = torch.randn(B, 3 * C, R, device=dev)
gr = torch.randn(B, C, T, device=dev)

def timed(fn, n=7):
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out

res = {}
res["sci_fwd"], u = timed(lambda: F_.sci(x, ks, rt))
res["cci_fwd"], o = timed(lambda: F_.cci(u, kc))
res["cci_sci_bwd"], _ = timed(lambda: torch.autograd.grad(o, (ks, kc), g, retain_graph=True))
res["rbf_fwd"], rec = timed(lambda: F_.rbf_readout(v, x, kr, rt))
res["rbf_bwd"], _ = timed(lambda: torch.autograd.grad(rec, (v, kr), gr, retain_graph=True))
scale = 1e6 / B
print(json.dumps({k: round(t * scale, 3) for k, t in res.items()} | {"B": B, "T": T, "R": R, "unit": "ms per 1M encounters"}))
