"""Build an in-memory scenario with the REFERENCE's own classes from synthetic tensors.

Used only by ``tests/golden/make_golden.py`` and oracle-pinning checks in the build container
(needs ``/root/reference``; see ``tools/ref_import.py``).  Recipe: SURVEY.md Appendix B step 2.
"""
from __future__ import annotations

import torch

from tools.ref_import import import_reference


def build_reference_scenario(ft: dict, number_of_rays: int, points_per_facet=(50, 50), device="cpu"):
    """``ft`` = ``artist_b200.scenario.synthetic.synthetic_field_tensors(...)``."""
    import_reference()
    from artist.field.heliostat_field import HeliostatField
    from artist.field.heliostat_group_rigid_body import HeliostatGroupRigidBody
    from artist.field.solar_tower import SolarTower
    from artist.field.tower_target_areas_cylindrical import TowerTargetAreasCylindrical
    from artist.field.tower_target_areas_planar import TowerTargetAreasPlanar
    from artist.nurbs.surfaces import NURBSSurfaces
    from artist.nurbs.utils import create_nurbs_evaluation_grid
    from artist.scenario.scenario import Scenario
    from artist.scene.light_source_array import LightSourceArray
    from artist.scene.sun import Sun

    dev = torch.device(device)
    n = ft["positions"].shape[0]
    f = ft["canting"].shape[1]
    grid = create_nurbs_evaluation_grid(torch.tensor(points_per_facet), device=dev)
    eval_points = grid[None, None].expand(n, f, -1, -1)
    surf = NURBSSurfaces(degrees=ft["nurbs_degrees"], control_points=ft["nurbs_control_points"], device=dev)
    pts, nrm = surf.calculate_surface_points_and_normals(
        evaluation_points=eval_points, canting=ft["canting"], facet_translations=ft["facet_translations"], device=dev)
    group = HeliostatGroupRigidBody(
        names=ft["names"], positions=ft["positions"], surface_points=pts.reshape(n, -1, 4),
        surface_normals=nrm.reshape(n, -1, 4), canting=ft["canting"], facet_translations=ft["facet_translations"],
        initial_orientations=ft["initial_orientations"], nurbs_control_points=ft["nurbs_control_points"],
        nurbs_degrees=ft["nurbs_degrees"],
        kinematics_translation_deviation_parameters=ft["translation_deviations"],
        kinematics_rotation_deviation_parameters=ft["rotation_deviations"],
        actuator_parameters_non_optimizable=ft["actuator_non_optimizable"],
        actuator_parameters_optimizable=ft["actuator_optimizable"], device=dev)
    field = HeliostatField([group], device=dev) if _takes_device(HeliostatField) else HeliostatField([group])
    planar = TowerTargetAreasPlanar(names=ft["planar_names"], centers=ft["planar_centers"],
                                    normals=ft["planar_normals"], dimensions=ft["planar_dimensions"])
    cyl = TowerTargetAreasCylindrical(names=ft["cyl_names"], centers=ft["cyl_centers"], normals=ft["cyl_normals"],
                                      axes=ft["cyl_axes"], radii=ft["cyl_radii"], heights=ft["cyl_heights"],
                                      opening_angles=ft["cyl_opening_angles"])
    tower = SolarTower([planar, cyl], device=dev)
    sun = Sun(number_of_rays=number_of_rays, device=dev)
    scenario = Scenario(power_plant_position=torch.tensor([50.91, 6.38, 87.0], dtype=torch.float64),
                        solar_tower=tower, light_sources=LightSourceArray([sun]), heliostat_field=field)
    return scenario, group, eval_points


def _takes_device(cls) -> bool:
    import inspect

    return "device" in inspect.signature(cls.__init__).parameters
