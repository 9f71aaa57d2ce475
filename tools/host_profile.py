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
st.sort_stats("cumulative").print_stats(int(os.environ.get("AB200_PROFILE_ROWS", "45")))

# ---- the same forward+backward captured once into a CUDA graph (tests/test_gpu_cuda_graph.py) ----
def fb():
    wl.param.grad = None
    total = wl.forward_local()
    (total * total).mean().backward()


wl.last_total = None   # (drop the last eager step's autograd graph: a live AccumulateGrad node of the default stream would
                       # invalidate the capture - torch's rule, not the kernels')
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3):
        fb()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    fb()
torch.cuda.synchronize()
eager = (time.perf_counter() - t0) / steps * 1e3
wl.param.grad = None
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    total = wl.forward_local()
    (total * total).mean().backward()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    graph.replay()
torch.cuda.synchronize()
replay = (time.perf_counter() - t0) / steps * 1e3
print(f"{n} heliostats, forward+backward: eager {eager:.3f} ms/step, CUDA-graph replay {replay:.3f} ms/step ({eager / replay:.1f}x)")
