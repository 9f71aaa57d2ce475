"""Host-side cost of one step at a small sample count (the launch-bound regime of BASELINE.json configs 2 and 4):
cProfile of N steps.  usage: python tools/host_profile.py [heliostats] [steps] [kind]"""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
kind = sys.argv[3] if len(sys.argv) > 3 else "surface"
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
wl = bench.Workload(dev, n, 1, 0, kind=kind)
for _ in range(20):
    wl.step()
torch.cuda.synchronize()
import time

t0 = time.perf_counter()
for _ in range(steps):
    wl.step()
torch.cuda.synchronize()
print(f"{n} heliostats: {(time.perf_counter() - t0) / steps * 1e3:.3f} ms/step wall")
pr = cProfile.Profile()
pr.enable()
for _ in range(steps):
    wl.step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
