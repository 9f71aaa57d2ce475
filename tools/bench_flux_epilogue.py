"""The steps after the tracer in the reference's optimisers at the bench size: centre-of-mass crop and the
pixel / KL-divergence losses on [N,256,256] bitmaps, forward + backward, kernel times and effective HBM bandwidth.
usage: python tools/bench_flux_epilogue.py [samples] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import __graft_entry__ as entry

entry.build()
from artist_b200 import _lib, build_synthetic_scenario  # noqa: E402
from artist_b200.flux import bitmap  # noqa: E402
from artist_b200.optim import KLDivergenceLoss, PixelLoss  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda:0")
scenario, group = build_synthetic_scenario(8, number_of_rays=2, points_per_facet=(4, 4), device=dev)
yy, xx = torch.meshgrid(torch.linspace(-1, 1, 256, device=dev), torch.linspace(-1, 1, 256, device=dev), indexing="ij")
flux = (torch.exp(-((xx - 0.1) ** 2 + (yy + 0.05) ** 2) / 0.02)[None] * (1 + 0.1 * torch.rand(n, 1, 1, device=dev))).contiguous()
truth = torch.exp(-(xx ** 2 + yy ** 2) / 0.03)[None].expand(n, -1, -1).contiguous()
tidx = torch.zeros(n, dtype=torch.int32, device=dev)
mb = n * 256 * 256 * 4 / 1e6


def run(loss_cls):
    x = flux.clone().requires_grad_(True)
    cropped = bitmap.crop_flux_distributions_around_center(x, scenario.solar_tower, tidx)
    loss = loss_cls()(prediction=cropped, ground_truth=truth, target_area_indices=tidx, reduction_dimensions=(1, 2), device=dev)
    loss.sum().backward()
    return loss


for cls in (PixelLoss, KLDivergenceLoss):
    for _ in range(3):
        run(cls)
    _lib.timing_enabled = True
    _lib.timing_events.clear()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run(cls)
    e1.record()
    torch.cuda.synchronize()
    _lib.timing_enabled = False
    kern = {k: sum(a.elapsed_time(b) for a, b in ev) / len(ev) for k, ev in sorted(_lib.timing_events.items())}
    print(f"{cls.__name__}: crop + loss forward+backward on [{n},256,256] ({mb:.0f} MB per bitmap stack): {e0.elapsed_time(e1) / reps:.3f} ms")
    for k, v in kern.items():
        print(f"    {k:28s} {v:7.3f} ms")
