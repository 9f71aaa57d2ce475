"""Small driver for ncu: a few full fwd+bwd steps of the bench workload at a reduced heliostat count."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 296
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
kind = sys.argv[3] if len(sys.argv) > 3 else "surface"       # "surface" | "motor" (bench.py --workload)
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
wl = bench.Workload(dev, n, 1, 0, kind=kind)
for _ in range(steps):
    wl.step()
torch.cuda.synchronize()
print("done", n, steps)
