"""Synthetic PAINT-like heliostat fields as plain tensors.

There is no network and no HDF5 library in the build/GPU images, so benchmarks and parity
tests use fields generated in memory: a ``ceil(sqrt(N))``-wide grid of identical four-facet
heliostats north of the tower, with the facet translations / canting vectors and the
linear-actuator constants of the Juelich heliostat AA39 (values as published in the PAINT
database entry the reference ships at ``tests/data/field_data/AA39-heliostat-properties.json``),
ideal planar NURBS control nets, one planar 8 m x 8 m target and the tilted Juelich
cylindrical ``receiver``.
"""
from __future__ import annotations

import math

import torch

# AA39 facet geometry (translation, canting_e, canting_n per facet).
_AA39_FACETS = [
    ([-0.8075, 0.6425, 0.0402], [0.8024845719337463, -0.0, -0.004984567873179913],
     [1.9569211872294545e-05, 0.6374921798706055, 0.0031505227088928223]),
    ([0.8075, 0.6425, 0.0402], [0.8024845719337463, -0.0, 0.004984567873179913],
     [-1.9569211872294545e-05, 0.6374921798706055, 0.0031505227088928223]),
    ([-0.8075, -0.6425, 0.0402], [0.8024845719337463, -0.0, -0.004984567873179913],
     [-1.9569211872294545e-05, 0.6374921798706055, -0.0031505227088928223]),
    ([0.8075, -0.6425, 0.0402], [0.8024845719337463, -0.0, 0.004984567873179913],
     [1.9569211872294545e-05, 0.6374921798706055, -0.0031505227088928223]),
]
# AA39 linear actuators: (clockwise, min, max, increment, offset, pivot radius, initial angle, initial stroke)
_AA39_ACTUATORS = [
    (0.0, 0.0, 68745.0, 154166.6667, 0.335308, 0.338095, 0.039009536, 0.077412795),
    (1.0, 0.0, 75308.0, 154166.6667, 0.340771, 0.3191, 0.943922248, 0.077522286),
]


def aa39_facets() -> tuple[torch.Tensor, torch.Tensor]:
    """Return ``(facet_translations[4,4], canting[4,2,4])`` (homogeneous w = 0)."""
    tr = torch.zeros(4, 4)
    cant = torch.zeros(4, 2, 4)
    for i, (t, ce, cn) in enumerate(_AA39_FACETS):
        tr[i, :3] = torch.tensor(t)
        cant[i, 0, :3] = torch.tensor(ce)
        cant[i, 1, :3] = torch.tensor(cn)
    return tr, cant


def aa39_linear_actuators(n: int) -> tuple[torch.Tensor, torch.Tensor]:
    """``(non_optimizable[n,7,2], optimizable[n,2,2])`` in the reference layout
    (``artist/util/indices.py``: type, clockwise, min, max, increment, offset, pivot radius /
    initial angle, initial stroke length).  Actuator one's initial angle carries the loader's
    south->up orientation fix of ``-pi/2`` (``artist/io/h5_scenario_parser.py:631-642``)."""
    non_opt = torch.zeros(n, 7, 2)
    opt = torch.zeros(n, 2, 2)
    for j, (cw, lo, hi, inc, off, rad, a0, s0) in enumerate(_AA39_ACTUATORS):
        non_opt[:, 0, j] = 0.0  # type: linear
        non_opt[:, 1, j] = cw
        non_opt[:, 2, j] = lo
        non_opt[:, 3, j] = hi
        non_opt[:, 4, j] = inc
        non_opt[:, 5, j] = off
        non_opt[:, 6, j] = rad
        opt[:, 0, j] = a0
        opt[:, 1, j] = s0
    opt[:, 0, 0] += -math.pi / 2
    return non_opt, opt


def planar_control_points(cu: int, cv: int, canting: torch.Tensor) -> torch.Tensor:
    """Flat control net per facet spanning +-|canting vector| (``artist/nurbs/utils.py:52-121``)."""
    f = canting.shape[0]
    cp = torch.zeros(f, cu, cv, 3, dtype=canting.dtype)
    dims = torch.norm(canting, dim=2)
    ul = torch.linspace(0, 1, cu, dtype=canting.dtype)
    vl = torch.linspace(0, 1, cv, dtype=canting.dtype)
    cp[..., 0] = (-dims[:, 0, None] + 2 * dims[:, 0, None] * ul)[:, :, None]
    cp[..., 1] = (-dims[:, 1, None] + 2 * dims[:, 1, None] * vl)[:, None, :]
    return cp


def synthetic_field_tensors(n_heliostats: int, control_points: tuple[int, int] = (10, 10),
                            degrees: tuple[int, int] = (3, 3), surface_bump: float = 0.0,
                            seed: int = 0, with_cylinder: bool = True, pitch: float = 5.0,
                            first_row_north: float = 30.0,
                            planar_center: tuple[float, float, float] = (0.0, 0.0, 50.0)) -> dict:
    """Plain CPU tensors describing the synthetic field.

    ``surface_bump`` > 0 perturbs the control-point heights (metres, deterministic by ``seed``)
    so that surfaces differ per heliostat (used by gradient tests).
    """
    n = n_heliostats
    s = max(1, math.ceil(math.sqrt(n)))
    idx = torch.arange(n)
    positions = torch.zeros(n, 4)
    positions[:, 0] = ((idx % s).float() - s / 2) * pitch
    positions[:, 1] = first_row_north + torch.div(idx, s, rounding_mode="floor").float() * pitch
    positions[:, 2] = 1.7
    positions[:, 3] = 1.0
    tr, cant = aa39_facets()
    cp = planar_control_points(control_points[0], control_points[1], cant)
    cp = cp[None].repeat(n, 1, 1, 1, 1).contiguous()
    if surface_bump > 0:
        g = torch.Generator().manual_seed(seed)
        cp[..., 2] += surface_bump * torch.randn(cp.shape[:-1], generator=g)
    non_opt, opt = aa39_linear_actuators(n)
    tdev = torch.zeros(n, 9)
    tdev[:, 7] = 0.175  # concentrator_translation_n
    rdev = torch.zeros(n, 4)
    out = dict(
        names=[f"H{i:05d}" for i in range(n)],
        positions=positions,
        facet_translations=tr[None].repeat(n, 1, 1).contiguous(),
        canting=cant[None].repeat(n, 1, 1, 1).contiguous(),
        nurbs_control_points=cp,
        nurbs_degrees=torch.tensor(degrees),
        initial_orientations=torch.tensor([0.0, 0.0, 1.0, 0.0])[None].repeat(n, 1),
        translation_deviations=tdev,
        rotation_deviations=rdev,
        actuator_non_optimizable=non_opt,
        actuator_optimizable=opt,
        planar_names=["receiver_plane"],
        planar_centers=torch.tensor([[planar_center[0], planar_center[1], planar_center[2], 1.0]]),
        planar_normals=torch.tensor([[0.0, 1.0, 0.0, 0.0]]),
        planar_dimensions=torch.tensor([[8.0, 8.0]]),
    )
    if with_cylinder:
        out.update(
            cyl_names=["receiver"],
            cyl_centers=torch.tensor([[0.0, -3.76, 56.7, 1.0]]),
            cyl_axes=torch.tensor([[0.0, 0.4226, 0.9063, 0.0]]),
            cyl_normals=torch.tensor([[0.0, 0.9063, -0.4226, 0.0]]),
            cyl_radii=torch.tensor([4.14]),
            cyl_heights=torch.tensor([5.229]),
            cyl_opening_angles=torch.tensor([1.0472]),
        )
    else:
        out.update(cyl_names=[], cyl_centers=torch.zeros(0, 4), cyl_axes=torch.zeros(0, 4),
                   cyl_normals=torch.zeros(0, 4), cyl_radii=torch.zeros(0), cyl_heights=torch.zeros(0),
                   cyl_opening_angles=torch.zeros(0))
    return out


def build_synthetic_scenario(n_heliostats: int, number_of_rays: int = 10, points_per_facet: tuple[int, int] = (50, 50),
                             control_points: tuple[int, int] = (10, 10), surface_bump: float = 0.0, seed: int = 0,
                             device="cuda", with_cylinder: bool = True, field_tensors: dict | None = None):
    """Assemble ``Scenario`` / ``HeliostatGroupRigidBody`` objects on ``device`` from the synthetic
    tensors (surface points and normals evaluated by the NURBS kernel).  Returns ``(scenario, group)``."""
    from ..field import (HeliostatField, HeliostatGroupRigidBody, SolarTower, TowerTargetAreasCylindrical,
                         TowerTargetAreasPlanar)
    from ..nurbs import NURBSSurfaces, create_nurbs_evaluation_grid
    from ..scene import LightSourceArray, Sun
    from .scenario import Scenario

    dev = torch.device(device)
    ft = field_tensors or synthetic_field_tensors(n_heliostats, control_points, surface_bump=surface_bump, seed=seed,
                                                  with_cylinder=with_cylinder)
    g = lambda k: ft[k].to(dev)
    n = ft["positions"].shape[0]
    grid = create_nurbs_evaluation_grid(torch.tensor(points_per_facet), device=dev)
    ev = grid[None, None].expand(n, ft["canting"].shape[1], -1, -1)
    surf = NURBSSurfaces(ft["nurbs_degrees"], g("nurbs_control_points"), device=dev)
    pts, nrm = surf.calculate_surface_points_and_normals(ev, g("canting"), g("facet_translations"))
    group = HeliostatGroupRigidBody(
        names=ft["names"], positions=g("positions"), surface_points=pts.reshape(n, -1, 4),
        surface_normals=nrm.reshape(n, -1, 4), canting=g("canting"), facet_translations=g("facet_translations"),
        initial_orientations=g("initial_orientations"), nurbs_control_points=g("nurbs_control_points"),
        nurbs_degrees=ft["nurbs_degrees"], kinematics_translation_deviation_parameters=g("translation_deviations"),
        kinematics_rotation_deviation_parameters=g("rotation_deviations"),
        actuator_parameters_non_optimizable=g("actuator_non_optimizable"),
        actuator_parameters_optimizable=g("actuator_optimizable"), device=dev)
    planar = TowerTargetAreasPlanar(ft["planar_names"], g("planar_centers"), g("planar_normals"), g("planar_dimensions"))
    cyl = TowerTargetAreasCylindrical(ft["cyl_names"], g("cyl_centers"), g("cyl_normals"), g("cyl_axes"),
                                      g("cyl_radii"), g("cyl_heights"), g("cyl_opening_angles"))
    scenario = Scenario(power_plant_position=torch.tensor([50.91, 6.38, 87.0], dtype=torch.float64),
                        solar_tower=SolarTower([planar, cyl], device=dev),
                        light_sources=LightSourceArray([Sun(number_of_rays=number_of_rays, device=dev)]),
                        heliostat_field=HeliostatField([group], device=dev))
    return scenario, group
