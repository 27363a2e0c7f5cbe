"""One forward call of the BiLSTM kernel (ncu target).   python benchmarks/_lstm_one.py [B] [R] [I]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_interpolation_clustering_b200.lstm import BiLSTMB200
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
R = int(sys.argv[2]) if len(sys.argv) > 2 else 96
I = int(sys.argv[3]) if len(sys.argv) > 3 else 18
dev = torch.device("cuda:0")
m = BiLSTMB200(I).to(dev)
x = torch.randn(R, B, I, device=dev)
with torch.no_grad():
    for _ in range(2):
        m(x)
torch.cuda.synchronize()
