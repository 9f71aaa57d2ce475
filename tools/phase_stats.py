"""GPU: where a trace CTA spends its time at the benchmark size - cycles of thread 0 per phase (AB200_PHASE in
csrc/trace.cu), summed over the CTAs of one forward and one backward launch, plus the window diagnostics."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from artist_b200 import ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
dev = torch.device("cuda:0")
wl = bench.Workload(dev, n, 1, 0)
for _ in range(2):
    wl.step()
torch.cuda.synchronize()
ops.trace_stats = torch.zeros(20, dtype=torch.int64, device=dev)
wl.step()
torch.cuda.synchronize()
st = ops.trace_stats.tolist()
ops.trace_stats = None
print(f"samples {n}: threads on the global fallback {st[0]}; mean window {st[1] / max(st[2], 1):.0f} cells")
for name, base, labels in (("forward", 4, ["start-up loads + window clear", "window placement", "clear bitmap row outside window",
                                            "ray loop (thread 0)", "wait for last warp", "window flush", "out-of-window tap conversion"]),
                           ("backward", 12, ["start-up loads", "window placement", "stage gradient window", "ray loop (thread 0)",
                                             "wait + dL/dO reduction"])):
    tot = sum(st[base:base + len(labels)])
    print(f"{name}: {tot / max(st[2], 1):.0f} cycles per CTA")
    for k, lab in enumerate(labels):
        print(f"  {lab:34s} {st[base + k] / max(st[2], 1):9.0f} cycles  {100.0 * st[base + k] / max(tot, 1):5.1f} %")
