"""GPU: bitmap-window diagnostics of the forward trace at the benchmark size (fallback use, window area)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from artist_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
dev = torch.device("cuda:0")
wl = bench.Workload(dev, n, 1, 0)
ops.trace_stats = torch.zeros(20, dtype=torch.int64, device=dev)
with torch.no_grad():
    flux, ic, ot, _ = wl.tracer.trace_rays(wl.inc, wl.mask, wl.tidx)
torch.cuda.synchronize()
st = ops.trace_stats.tolist()
print(f"samples {n}: threads that used the global fallback {st[0]} of {st[2] * 1024 if st[2] else 0}; mean window {st[1] / max(st[2], 1):.0f} cells "
      f"(capacity {os.environ.get('AB200_WIN_KB', '200')} KB = {int(os.environ.get('AB200_WIN_KB', '200')) * 256} cells); mean intercept factor {ic.mean().item():.3f}")
