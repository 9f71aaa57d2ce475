"""Generate ``tests/golden/hotpath_golden.pt`` by running the REAL reference (imported from
``/root/reference``, build container only) on small seeded inputs.  The committed fixture pins the
oracle (``tests/test_oracle_golden.py``) and is also compared with the CUDA path directly
(``tests/test_gpu_golden.py``).  Re-run:  python tests/golden/make_golden.py
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from artist_b200.scenario.synthetic import synthetic_field_tensors  # noqa: E402
from tools.ref_import import import_reference  # noqa: E402
from tools.ref_scenario import build_reference_scenario  # noqa: E402

CPU = torch.device("cpu")


def trace_case(ideal: bool, seed: int):
    """4 heliostats, targets planar/cylindrical alternating, full reference pipeline on the CPU."""
    from artist.raytracing import geometry
    from artist.raytracing.heliostat_ray_tracer import HeliostatRayTracer
    from artist.scene.rays import Rays

    n, rays, ppf, res = 4, 4, (8, 8), (32, 24)
    ft = synthetic_field_tensors(n, control_points=(6, 6), surface_bump=0.003, seed=seed)
    g = torch.Generator().manual_seed(100 + seed)
    ft["rotation_deviations"] = 0.01 * torch.randn(n, 4, generator=g)
    ft["translation_deviations"][:, :6] = 0.01 * torch.randn(n, 6, generator=g)
    if ideal:
        ft["actuator_non_optimizable"][:, 0] = 1.0
        ft["actuator_non_optimizable"][:, 2] = -10.0
        ft["actuator_non_optimizable"][:, 3] = 10.0
    scenario, group, evalp = build_reference_scenario(ft, rays, ppf)
    mask = torch.ones(n, dtype=torch.int32)
    tidx = torch.tensor([0, 1, 0, 1], dtype=torch.int32)
    inc = torch.nn.functional.normalize(torch.tensor([[0.0, 1.0, -0.3], [0.2, 0.9, -0.3], [-0.1, 1.0, -0.2], [0.3, 0.8, -0.4]]), dim=1)
    inc = torch.cat([inc, torch.zeros(n, 1)], 1)
    group.activate_heliostats(mask, device=CPU)
    aim = scenario.solar_tower.get_centers_of_target_areas(tidx, device=CPU)
    group.align_surfaces_with_incident_ray_directions(aim, inc, mask, device=CPU)
    motor = group.kinematics.active_motor_positions.clone()
    ori_incident = group.kinematics.incident_ray_directions_to_orientations(inc, aim, device=CPU).clone()
    ori_motor = group.kinematics.motor_positions_to_orientations(motor, device=CPU).clone()
    tracer = HeliostatRayTracer(scenario, group, blocking_active=False, batch_size=3, random_seed=7,
                                bitmap_resolution=torch.tensor(res))
    flux, ic, ot, bl = tracer.trace_rays(inc, mask, tidx, device=CPU)
    per_target = tracer.get_bitmaps_per_target(flux, tidx, device=CPU)
    du, de = tracer.distortions_dataset.distortions_u, tracer.distortions_dataset.distortions_e
    # per-ray intermediates from the reference's free functions
    ref_dirs = geometry.reflect(inc.unsqueeze(1), group.active_surface_normals)
    scattered = tracer.scatter_rays(du, de, ref_dirs, device=CPU)
    planar = tidx < 1
    be = torch.zeros(n, rays, group.active_surface_points.shape[1])
    bu, tt, lam = torch.zeros_like(be), torch.zeros_like(be), torch.zeros_like(be)
    be[planar], bu[planar], tt[planar], lam[planar] = geometry.line_plane_intersections(
        Rays(scattered.ray_directions[planar], scattered.ray_magnitudes[planar]), group.active_surface_points[planar],
        scenario.solar_tower.target_areas[0], tidx[planar], torch.tensor(res), device=CPU)
    be[~planar], bu[~planar], tt[~planar], lam[~planar] = geometry.line_cylinder_intersections(
        Rays(scattered.ray_directions[~planar], scattered.ray_magnitudes[~planar]), group.active_surface_points[~planar],
        scenario.solar_tower.target_areas[1], tidx[~planar] - 1, torch.tensor(res), device=CPU)
    return dict(
        field={k: v for k, v in ft.items() if isinstance(v, torch.Tensor)}, rays=rays, points_per_facet=ppf, res=res,
        target_idx=tidx, incident=inc, aim=aim, surface_points=group.surface_points.clone(),
        surface_normals=group.surface_normals.clone(), motor=motor, orientations_incident=ori_incident,
        orientations_motor=ori_motor, aligned_points=group.active_surface_points.clone(),
        aligned_normals=group.active_surface_normals.clone(), dist_u=du.clone(), dist_e=de.clone(),
        reflected=ref_dirs.clone(), scattered=scattered.ray_directions.clone(), be=be, bu=bu, t=tt, lambert=lam,
        flux=flux.clone(), intercept=ic.clone(), on_target=ot.clone(), blocking=bl.clone(), per_target=per_target.clone())


def blocking_case():
    """3x3 heliostats at 3.4 m pitch aiming at a low target: the reference with blocking_active=True (two batch sizes)."""
    from artist.raytracing import blocking as ref_blocking
    from artist.raytracing.heliostat_ray_tracer import HeliostatRayTracer

    n, rays, ppf, res = 9, 4, (10, 10), (48, 48)
    ft = synthetic_field_tensors(n, control_points=(6, 6), surface_bump=0.001, pitch=3.4, planar_center=(0.0, 0.0, 6.0))
    scenario, group, _ = build_reference_scenario(ft, rays, ppf)
    mask, tidx, inc = scenario.index_mapping(group, single_incident_ray_direction=torch.tensor([0.0, 0.8, -0.6, 0.0]), device=CPU)
    group.activate_heliostats(mask, device=CPU)
    aim = scenario.solar_tower.get_centers_of_target_areas(tidx, device=CPU)
    group.align_surfaces_with_incident_ray_directions(aim, inc, mask, device=CPU)
    out = dict(field={k: v for k, v in ft.items() if isinstance(v, torch.Tensor)}, rays=rays, res=res, target_idx=tidx,
               incident=inc, aligned_points=group.active_surface_points.clone(), aligned_normals=group.active_surface_normals.clone())
    for bs in (100, 4):
        tracer = HeliostatRayTracer(scenario, group, blocking_active=True, batch_size=bs, random_seed=7, bitmap_resolution=torch.tensor(res))
        flux, ic, ot, bl = tracer.trace_rays(inc, mask, tidx, device=CPU)
        out[f"batch{bs}"] = dict(flux=flux.clone(), intercept=ic.clone(), on_target=ot.clone(), blocking=bl.clone())
    c, s, nn = ref_blocking.create_blocking_primitives_rectangles_by_index(tracer.blocking_heliostat_surfaces_active, device=CPU)
    out.update(corners=c.clone(), spans=s.clone(), normals=nn.clone(), dist_u=tracer.distortions_dataset.distortions_u.clone(),
               dist_e=tracer.distortions_dataset.distortions_e.clone())
    return out


def nurbs_case(degrees, cps, ppf, canting: bool, seed: int):
    from artist.nurbs.surfaces import NURBSSurfaces
    from artist.nurbs.utils import create_nurbs_evaluation_grid

    n = 2
    ft = synthetic_field_tensors(n, control_points=cps, surface_bump=0.004, seed=seed)
    ev = create_nurbs_evaluation_grid(torch.tensor(ppf), device=CPU)[None, None].expand(n, 4, -1, -1)
    cp = ft["nurbs_control_points"].clone().requires_grad_(True)
    surf = NURBSSurfaces(torch.tensor(degrees), cp, device=CPU)
    pts, nrm = surf.calculate_surface_points_and_normals(ev, ft["canting"] if canting else None,
                                                         ft["facet_translations"] if canting else None, device=CPU)
    g = torch.Generator().manual_seed(seed)
    wp, wn = torch.randn(pts.shape, generator=g), torch.randn(nrm.shape, generator=g)
    ((pts * wp).sum() + (nrm * wn).sum()).backward()
    return dict(degrees=degrees, control_points=ft["nurbs_control_points"], canting=ft["canting"] if canting else None,
                facet_translations=ft["facet_translations"] if canting else None, eval_points=ev.contiguous().clone(),
                points=pts.detach().clone(), normals=nrm.detach().clone(), weight_points=wp, weight_normals=wn,
                grad_control_points=cp.grad.clone())


def main() -> None:
    import_reference()
    from artist.raytracing.sampling import RestrictedDistributedSampler
    from artist.scene.sun import Sun

    out = {"trace_linear": trace_case(False, 1), "trace_ideal": trace_case(True, 2), "blocking": blocking_case(),
           "nurbs": [nurbs_case((3, 3), (6, 7), (7, 5), True, 1), nurbs_case((3, 3), (10, 10), (9, 9), False, 2),
                     nurbs_case((2, 3), (5, 6), (6, 6), True, 3)]}
    samplers = {}
    for ns, nh, ws in [(12, 4, 1), (12, 4, 2), (12, 4, 3), (12, 4, 4), (4, 1, 3), (4, 2, 3), (24, 6, 4), (7, 7, 8)]:
        samplers[(ns, nh, ws)] = [list(RestrictedDistributedSampler(ns, nh, ws, r)) for r in range(ws)]
    out["samplers"] = samplers
    du, de = Sun(number_of_rays=3, device=CPU).get_distortions(number_of_points=5, number_of_active_heliostats=2, random_seed=7)
    out["sun_seed7"] = dict(u=du.clone(), e=de.clone(), next_rand=torch.rand(3))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hotpath_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes; torch", torch.__version__)


if __name__ == "__main__":
    main()
