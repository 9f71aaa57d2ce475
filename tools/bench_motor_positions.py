"""BASELINE.json configs[4]: motor-position optimisation step over the full synthetic field - gradients of the flux loss
to the motor positions through the fused trace (dL/dO reduced per sample inside trace_bwd) and the kinematics backward
kernel.  Prints ms per step and the kernel times (not the headline metric; bench.py measures that).
usage: python tools/bench_motor_positions.py [heliostats] [steps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import __graft_entry__ as entry

entry.build()
from artist_b200 import HeliostatRayTracer, _lib, build_synthetic_scenario  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda:0")
scenario, group = build_synthetic_scenario(n, number_of_rays=10, points_per_facet=(50, 50), surface_bump=1e-4, device=dev)
mask, tidx, inc = scenario.index_mapping(group)
aim = scenario.solar_tower.get_centers_of_target_areas(tidx)
group.activate_heliostats(mask)
group.align_surfaces_with_incident_ray_directions(aim, inc, mask)
motor = group.kinematics.active_motor_positions.detach().clone().requires_grad_(True)
opt = torch.optim.Adam([motor], lr=1.0, fused=True)
tracer = HeliostatRayTracer(scenario, group, blocking_active=False, bitmap_resolution=torch.tensor([256, 256]))


def step():
    opt.zero_grad(set_to_none=True)
    group.activate_heliostats(mask)
    group.align_surfaces_with_motor_positions(motor, mask)
    flux, *_ = tracer.trace_rays(inc, mask, tidx)
    total = tracer.get_bitmaps_per_target(flux, tidx)
    loss = (total * total).mean()
    loss.backward()
    opt.step()
    return loss


for _ in range(3):
    step()
_lib.timing_enabled = True
_lib.timing_events.clear()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
kern = {k: round(sum(a.elapsed_time(b) for a, b in ev) / len(ev), 4) for k, ev in sorted(_lib.timing_events.items())}
rays = n * group.surface_points.shape[1] * 10
print(f"motor-position step: {ms:.3f} ms, {rays / ms * 1e3:.3e} rays/s, loss {float(loss):.4e}, grad max {float(motor.grad.abs().max()):.3e}")
print(kern)
