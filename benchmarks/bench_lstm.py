"""BiLSTM (hidden 128, R steps) on the persistent tcgen05 kernel vs cuDNN through torch.nn.LSTM, forward and
forward + backward, at the encoder (I = 18) and decoder (I = 256) widths.    python benchmarks/bench_lstm.py [B] [R]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_interpolation_clustering_b200.lstm import BiLSTMB200

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
R = int(sys.argv[2]) if len(sys.argv) > 2 else 96
dev = torch.device("cuda:0")


def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


res = {"B": B, "R": R, "unit": "ms per call (best of 3)"}
for I in (18, 256):
    ours = BiLSTMB200(I).to(dev)
    ref = torch.nn.LSTM(I, 128, bidirectional=True).to(dev)
    ref.load_state_dict(ours.state_dict())
    x = torch.randn(R, B, I, device=dev)
    g = torch.randn(R, B, 256, device=dev)
    xr = x.clone().requires_grad_(True)

    def fwd(m):
        with torch.no_grad():
            return m(x)

    def fwdbwd(m):
        for p in m.parameters():
            p.grad = None
        xr.grad = None
        out, _ = m(xr)
        out.backward(g)

    r = {}
    r["ours_fwd"] = timed(lambda: fwd(ours))
    r["ours_fwd_bwd"] = timed(lambda: fwdbwd(ours))
    for tf32 in (True, False):
        torch.backends.cudnn.allow_tf32 = tf32
        tag = "cudnn_tf32" if tf32 else "cudnn_fp32"
        r[tag + "_fwd"] = timed(lambda: fwd(ref))
        r[tag + "_fwd_bwd"] = timed(lambda: fwdbwd(ref))
    with torch.no_grad():
        a = ours(x)[0]
        torch.backends.cudnn.allow_tf32 = False
        b = ref(x)[0]
        torch.backends.cudnn.allow_tf32 = True
        c = ref(x)[0]
    r["max_abs_diff_vs_cudnn_fp32"] = float((a - b).abs().max())
    r["max_abs_diff_cudnn_tf32_vs_fp32"] = float((c - b).abs().max())
    r["ours_fwd_encounters_per_s"] = round(B / (r["ours_fwd"] * 1e-3))
    res[f"I{I}"] = {k: (round(v, 3) if isinstance(v, float) and v > 1e-3 else v) for k, v in r.items()}
print(json.dumps(res))
