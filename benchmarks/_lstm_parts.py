"""Where one BiLSTM forward call spends its time: glue (weight permutation), the input-projection GEMM, the weight
packing and the persistent recurrence kernel.   python benchmarks/_lstm_parts.py [B] [R] [I]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_interpolation_clustering_b200 import _lib
from deep_interpolation_clustering_b200.lstm import BiLSTMB200, _perm_index, H
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
R = int(sys.argv[2]) if len(sys.argv) > 2 else 96
dev = torch.device("cuda:0")
L = _lib.lib()

def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return round(best, 3), r

out_all = {"B": B, "R": R}
for I in (18, 256):
    m = BiLSTMB200(I).to(dev)
    x = torch.randn(R, B, I, device=dev)
    res = {}
    with torch.no_grad():
        res["perm_index"], perm = timed(lambda: _perm_index(dev))
        res["permute_weights"], (wp, bp) = timed(lambda: (torch.cat([m.weight_ih_l0, m.weight_ih_l0_reverse], 0)[perm],
                                                          torch.cat([m.bias_ih_l0 + m.bias_hh_l0, m.bias_ih_l0_reverse + m.bias_hh_l0_reverse], 0)[perm]))
        res["addmm_f32_simt"], pre = timed(lambda: torch.addmm(bp, x.reshape(R * B, I), wp.t()))
        st = _lib.current_stream(dev)
        wpc, bpc, x2 = wp.contiguous(), bp.contiguous(), x.reshape(R * B, I)
        pw = torch.empty(int(L.dic_lstm_project_packed_bytes(I)), dtype=torch.uint8, device=dev)
        res["pack_wih"], _ = timed(lambda: L.dic_lstm_pack_wih(wpc.data_ptr(), pw.data_ptr(), I, st))
        pre2 = torch.empty((R * B, 1024), device=dev)
        res["project_tcgen05"], _ = timed(lambda: L.dic_lstm_project(x2.data_ptr(), I, pw.data_ptr(), bpc.data_ptr(), pre2.data_ptr(), R * B, I, 0, st))
        res["project_max_abs_diff_vs_f32"] = float((pre2 - pre).abs().max())
        del pre2
        packed = torch.empty(int(L.dic_lstm_packed_bytes()), dtype=torch.uint8, device=dev)
        res["pack"], _ = timed(lambda: L.dic_lstm_pack_whh(m.weight_hh_l0.data_ptr(), m.weight_hh_l0_reverse.data_ptr(), packed.data_ptr(), st))
        out = torch.empty((R, B, 2 * H), device=dev); hn = torch.empty((2, B, H), device=dev); cn = torch.empty((2, B, H), device=dev)
        res["kernel_inference"], _ = timed(lambda: L.dic_lstm_fwd(pre.data_ptr(), packed.data_ptr(), None, None, out.data_ptr(), hn.data_ptr(), cn.data_ptr(), None, R, B, H, st))
        save = torch.empty((2, R, B, 5, H), device=dev)
        res["kernel_training"], _ = timed(lambda: L.dic_lstm_fwd(pre.data_ptr(), packed.data_ptr(), None, None, out.data_ptr(), hn.data_ptr(), cn.data_ptr(), save.data_ptr(), R, B, H, st))
        res["whole_forward"], _ = timed(lambda: m(x))
        del save, pre
    out_all[f"I{I}"] = res
print(json.dumps(out_all))
