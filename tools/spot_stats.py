"""GPU: focal-spot bounding boxes of the benchmark field (undistorted reflections), to size the bitmap window."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from artist_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
dev = torch.device("cuda:0")
wl = bench.Workload(dev, n, 1, 0)
g = wl.group
zeros = torch.zeros(n, 1, wl.p, 2, device=dev)
opt = ops.TraceOptions(res_e=256, res_u=256, scatter_sigma=2.09e-3)
_, (be, bu, t, lam) = ops.trace_debug(g.active_surface_points, g.active_surface_normals, wl.inc, zeros, wl.tidx, wl.tracer._targets, opt)
valid = lam[:, 0] > 0
big = torch.tensor(1e9, device=dev)
bemin = torch.where(valid, be[:, 0], big).min(1).values; bemax = torch.where(valid, be[:, 0], -big).max(1).values
bumin = torch.where(valid, bu[:, 0], big).min(1).values; bumax = torch.where(valid, bu[:, 0], -big).max(1).values
w, h = (bemax - bemin).clamp_min(0), (bumax - bumin).clamp_min(0)
dist = t[:, 0].max(1).values
sig_px = 2.09e-3 * dist * 255 / 8
q = torch.tensor([0.05, 0.25, 0.5, 0.75, 0.95], device=dev)
print("valid centre-ray fraction", valid.float().mean().item())
print("distance  quantiles", torch.quantile(dist, q).tolist())
print("bbox width quantiles (px)", torch.quantile(w, q).tolist())
print("bbox height quantiles (px)", torch.quantile(h, q).tolist())
print("sigma (px) quantiles", torch.quantile(sig_px, q).tolist())
for k in (3.0, 4.0):
    area = (w + 2 * k * sig_px + 4).clamp(max=256) * (h + 2 * k * sig_px + 4).clamp(max=256)
    print(f"window cells at {k} sigma: quantiles", torch.quantile(area, q).tolist(), "fraction > 51200:", (area > 51200).float().mean().item())
