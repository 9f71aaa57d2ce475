"""Seeded synthetic cases shared by the parity tests: inputs are produced on the CPU with the
oracle (NURBS evaluation + alignment), then traced by the oracle (checker) and by the CUDA path."""
from __future__ import annotations

import torch

from artist_b200.scenario.synthetic import synthetic_field_tensors
from oracle import artist_oracle as O


targets_from = O.targets_from_field_tensors


def incident_directions(n: int, seed: int = 3, spread: float = 0.25) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    d = torch.tensor([0.0, 1.0, -0.35]) + spread * (torch.rand(n, 3, generator=g) - 0.5)
    d = torch.nn.functional.normalize(d, dim=1)
    return torch.cat([d, torch.zeros(n, 1)], dim=1)


aim_points = O.aim_points


def make_case(n: int = 4, points_per_facet=(12, 12), rays: int = 5, control_points=(6, 6), bump: float = 0.002,
              target_pattern=(0,), seed: int = 7, field_seed: int = 0):
    """Returns a dict with everything needed by the oracle and the CUDA ops (CPU tensors)."""
    ft = synthetic_field_tensors(n, control_points=control_points, surface_bump=bump, seed=field_seed)
    g = torch.Generator().manual_seed(field_seed + 11)
    ft["rotation_deviations"] = 0.01 * torch.randn(n, 4, generator=g)
    ft["translation_deviations"][:, :6] = 0.01 * torch.randn(n, 6, generator=g)
    tg = targets_from(ft)
    ev = O.nurbs_evaluation_grid(*points_per_facet)[None, None].expand(n, 4, -1, -1)
    pts, nrm = O.nurbs_points_and_normals(ft["nurbs_control_points"], 3, 3, ev, ft["canting"], ft["facet_translations"])
    pts, nrm = pts.reshape(n, -1, 4), nrm.reshape(n, -1, 4)
    tidx = torch.tensor([target_pattern[i % len(target_pattern)] for i in range(n)], dtype=torch.int32)
    inc = incident_directions(n, seed=seed)
    kin = O.Kin(ft["positions"], ft["translation_deviations"], ft["rotation_deviations"],
                ft["actuator_non_optimizable"], ft["actuator_optimizable"], True)
    aim = aim_points(tg, tidx)
    ori, motor = O.incident_ray_directions_to_orientations(kin, inc, aim)
    ap, an = O.align_surfaces(pts, nrm, ori)
    du, de = O.sun_distortions(rays, ap.shape[1], n, seed)
    return dict(ft=ft, targets=tg, eval_points=ev, surface_points=pts, surface_normals=nrm, target_idx=tidx,
                incident=inc, kin=kin, aim=aim, orientations=ori, motor=motor, points=ap, normals=an,
                dist_u=du, dist_e=de, rays=rays)


def cpu_trig(du: torch.Tensor, de: torch.Tensor) -> torch.Tensor:
    """[N,R,P,4] = cos u, sin u, cos e, sin e computed by torch on the CPU (strict-parity table)."""
    return torch.stack([torch.cos(du), torch.sin(du), torch.cos(de), torch.sin(de)], dim=-1).contiguous()


def oracle_trace_with_grads(case, res, wgt, dtype=torch.float32):
    """Oracle forward + autograd to (points, normals) in ``dtype`` (float64 = conditioning-free gold)."""
    old = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        cv = lambda x: x.to(dtype) if x.is_floating_point() else x
        tg = case["targets"]
        tgt = O.Targets(*[cv(getattr(tg, f)) for f in ("planar_centers", "planar_normals", "planar_dimensions",
                                                       "cyl_centers", "cyl_normals", "cyl_axes", "cyl_radii",
                                                       "cyl_heights", "cyl_opening_angles")])
        p = cv(case["points"]).clone().requires_grad_(True)
        n = cv(case["normals"]).clone().requires_grad_(True)
        flux, *_ = O.trace_rays(p, n, cv(case["incident"]), cv(case["dist_u"]), cv(case["dist_e"]), case["target_idx"],
                                tgt, res)
        (flux * cv(wgt)).sum().backward()
        return flux.detach(), p.grad, n.grad
    finally:
        torch.set_default_dtype(old)
