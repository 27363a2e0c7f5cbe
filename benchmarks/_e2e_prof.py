"""Kernel-level breakdown of one e2e step of bench.py (torch.profiler, CUDA activities)."""
import sys, types, json
sys.path.insert(0, '.')
import torch
import bench
from torch.profiler import profile, ProfilerActivity

args = types.SimpleNamespace(e2e_chunk=32768, e2e_encounters=131072, steps=4, e2e_upload="packed")
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
hp = bench.HotPath(131072, dev, seed=0)
# run the arm once for warm-up, then under the profiler
bench.e2e_arm(args, hp, dev, 1)
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    r = bench.e2e_arm(args, hp, dev, 1)
print(json.dumps(r))
ka = prof.key_averages()
rows = sorted(((k.key, k.device_time_total, k.count) for k in ka if k.device_time_total > 0), key=lambda t: -t[1])
tot = sum(t[1] for t in rows)
for name, us, n in rows[:40]:
    print(f"{us/1e3:10.3f} ms {100*us/tot:5.1f}% x{n:5d}  {name[:110]}")
