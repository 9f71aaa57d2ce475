"""Forward trace only at the bench size (tuning: A/B of library variants through AB200_LIB)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
dev = torch.device("cuda:0")
wl = bench.Workload(dev, 2048, 1, 0)
g = wl.group
with torch.no_grad():
    g.activate_heliostats(wl.mask)
    pts, nrm = wl.surf.calculate_surface_points_and_normals(wl.ev, g.active_canting, g.active_facet_translations)
    g.active_surface_points = pts.reshape(wl.n, -1, 4); g.active_surface_normals = nrm.reshape(wl.n, -1, 4)
    g.align_surfaces_with_incident_ray_directions(wl.aim, wl.inc, wl.mask)
    for _ in range(150):
        wl.tracer.trace_rays(wl.inc, wl.mask, wl.tidx)
    torch.cuda.synchronize()
    best = []
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            out = wl.tracer.trace_rays(wl.inc, wl.mask, wl.tidx)
        e1.record(); torch.cuda.synchronize()
        best.append(e0.elapsed_time(e1) / 50)
print(os.environ.get("AB200_LIB", "default"), "trace forward ms per launch, 6 batches of 50:", " ".join(f"{b:.4f}" for b in best), float(out[0].sum()))
