"""Config-5 diagnosis: motor-position step with the four combinations of (planar | cylindrical target) x (blocking off | on);
kernel times and the blocking candidate-count distribution.  usage: python tools/diag_motor.py [heliostats] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import __graft_entry__ as entry

entry.build()
from artist_b200 import HeliostatRayTracer, _lib, build_synthetic_scenario, ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda:0")
scenario, group = build_synthetic_scenario(n, number_of_rays=10, points_per_facet=(50, 50), surface_bump=1e-4, device=dev)
for target in (0, 1):
    for blocking in (False, True):
        mask, tidx, inc = scenario.index_mapping(group, single_target_area_index=target)
        aim = scenario.solar_tower.get_centers_of_target_areas(tidx)
        group.activate_heliostats(mask)
        group.align_surfaces_with_incident_ray_directions(aim, inc, mask)
        motor = group.kinematics.active_motor_positions.detach().clone().requires_grad_(True)
        tracer = HeliostatRayTracer(scenario, group, blocking_active=blocking, bitmap_resolution=torch.tensor([256, 256]))

        def step():
            motor.grad = None
            group.activate_heliostats(mask)
            group.align_surfaces_with_motor_positions(motor, mask)
            flux, ic, ot, bl = tracer.trace_rays(inc, mask, tidx)
            total = tracer.get_bitmaps_per_target(flux, tidx)
            (total * total).mean().backward()
            return flux, ic, ot, bl

        for _ in range(2):
            out = step()
        if blocking:
            ops.trace_stats = torch.zeros(20, dtype=torch.int64, device=dev)
            step()
            st = ops.trace_stats.tolist()
            ops.trace_stats = None
            print(f"    deferred points {st[17] / (n * 10000):.3f} of all; fwd cycles/CTA pass1 {st[7] / n:.0f} pass2 {st[18] / n:.0f}; "
                  f"bwd pass1 {st[15] / n:.0f} pass2 {st[19] / n:.0f}")
        _lib.timing_enabled = True
        _lib.timing_events.clear()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = step()
        e1.record()
        torch.cuda.synchronize()
        _lib.timing_enabled = False
        kern = {k: round(sum(a.elapsed_time(b) for a, b in ev) / len(ev), 3) for k, ev in sorted(_lib.timing_events.items())}
        print(f"target {'cylinder' if target else 'planar'} blocking {blocking}: {e0.elapsed_time(e1) / steps:.3f} ms/step; "
              f"intercept {float(out[1].mean()):.3f} on_target {float(out[2].mean()):.3f} unblocked {float(out[3].mean()):.3f}")
        print("   ", kern)
        if blocking:
            bi = tracer._blocking_inputs(tidx)
            opt = ops.TraceOptions(scatter_sigma=getattr(tracer.light_source, "scatter_sigma", 0.0))
            prims, cand_idx, cand_count = ops._prepare_blocking(bi, opt, n, dev)
            c = cand_count.float()
            print(f"    candidates per sample: mean {float(c.mean()):.1f} median {float(c.median()):.0f} max {int(c.max())} "
                  f"zero {int((c == 0).sum())}; overflow {int(ops.last_blocking_overflow)}")
