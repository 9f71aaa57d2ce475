from __future__ import annotations

from collections import defaultdict

import torch

from ..util.env import get_device


class Scenario:
    """Scenario container + ``index_mapping`` (``artist/scenario/scenario.py:25-470``)."""

    def __init__(self, power_plant_position: torch.Tensor, solar_tower, light_sources, heliostat_field) -> None:
        self.power_plant_position = power_plant_position
        self.solar_tower = solar_tower
        self.light_sources = light_sources
        self.heliostat_field = heliostat_field

    # ---- HDF5 scenario files (``:84-259``) -----------------------------------------------------------------------
    @staticmethod
    def get_number_of_heliostat_groups_from_hdf5(scenario_path) -> int:
        """``number_of_heliostat_groups`` of a scenario file (``:84-102``); read with the built-in HDF5 reader."""
        from ..io import h5lite

        with h5lite.File(scenario_path) as f:
            return int(f["number_of_heliostat_groups"][()])

    @classmethod
    def load_scenario_from_hdf5(cls, scenario_file, number_of_surface_points_per_facet: torch.Tensor = torch.tensor([50, 50]),
                                change_number_of_control_points_per_facet: torch.Tensor | None = None,
                                device: torch.device | None = None) -> "Scenario":
        """Load a scenario (``:105-259``).  ``scenario_file`` is an open ``artist_b200.io.File`` (or ``h5py.File``),
        or a path.  Heliostats are grouped by ``"<kinematics>_<actuator>"`` type exactly as the reference does; each
        group's surface points / normals are evaluated from its NURBS control nets by ``ab200_nurbs_fwd`` in one launch
        (the reference evaluates every heliostat twice in a Python loop, ``heliostat_field.py:318-335``)."""
        from ..field import (HeliostatField, HeliostatGroupRigidBody, SolarTower, TowerTargetAreasCylindrical,
                             TowerTargetAreasPlanar)
        from ..io import h5_scenario_parser, h5lite
        from ..nurbs import NURBSSurfaces, create_nurbs_evaluation_grid, create_planar_nurbs_control_points
        from ..scene import LightSourceArray, Sun

        dev = get_device(device)
        if isinstance(scenario_file, (str, bytes)) or hasattr(scenario_file, "__fspath__"):
            with h5lite.File(scenario_file) as f:
                parsed = h5_scenario_parser.parse_scenario(f)
        else:
            parsed = h5_scenario_parser.parse_scenario(scenario_file)
        tg = parsed["targets"]
        d = lambda t: t.to(dev)
        planar = TowerTargetAreasPlanar(tg["planar_names"], d(tg["planar_centers"]), d(tg["planar_normals"]),
                                        d(tg["planar_dimensions"]))
        cyl = TowerTargetAreasCylindrical(tg["cyl_names"], d(tg["cyl_centers"]), d(tg["cyl_normals"]), d(tg["cyl_axes"]),
                                          d(tg["cyl_radii"]), d(tg["cyl_heights"]), d(tg["cyl_opening_angles"]))
        suns = [Sun(number_of_rays=ls["number_of_rays"], distribution_parameters=ls["distribution_parameters"], device=dev)
                for ls in parsed["light_sources"]]
        per_facet = [int(v) for v in number_of_surface_points_per_facet.tolist()]
        grid = create_nurbs_evaluation_grid(torch.tensor(per_facet), device=dev)
        groups = []
        for key, g in parsed["groups"].items():
            if g["kinematics_type"] != "rigid_body":
                raise KeyError(f"heliostat group type {key} is not supported")
            n, n_facets = len(g["names"]), g["canting"].shape[1]
            control_points, canting = d(g["nurbs_control_points"]), d(g["canting"])
            if change_number_of_control_points_per_facet is not None:
                control_points = torch.stack([create_planar_nurbs_control_points(
                    number_of_control_points=change_number_of_control_points_per_facet, canting=canting[i], device=dev)
                    for i in range(n)])
            # a flat control net (all heights 0: ideal surface) gets canted and translated, a fitted one already
            # carries both (``artist/field/surface.py:99-120``) - decided per heliostat, evaluated in <= 2 launches
            translations = d(g["facet_translations"])
            flat = (control_points[..., 2] == 0).flatten(1).all(dim=1)
            pts = torch.empty(n, n_facets, grid.shape[0], 4, device=dev)
            nrm = torch.empty_like(pts)
            with torch.no_grad():
                for rows, cant in ((flat.nonzero().flatten(), True), ((~flat).nonzero().flatten(), False)):
                    if rows.numel() == 0:
                        continue
                    surfaces = NURBSSurfaces(g["nurbs_degrees"], control_points[rows].contiguous(), device=dev)
                    ev = grid[None, None].expand(rows.numel(), n_facets, -1, -1)
                    a, b = surfaces.calculate_surface_points_and_normals(
                        ev, canting[rows].contiguous() if cant else None, translations[rows].contiguous() if cant else None)
                    pts[rows], nrm[rows] = a, b
            groups.append(HeliostatGroupRigidBody(
                names=g["names"], positions=d(g["positions"]), surface_points=pts.reshape(n, -1, 4),
                surface_normals=nrm.reshape(n, -1, 4), canting=canting, facet_translations=translations,
                initial_orientations=d(g["initial_orientations"]), nurbs_control_points=control_points,
                nurbs_degrees=g["nurbs_degrees"], kinematics_translation_deviation_parameters=d(g["translation_deviations"]),
                kinematics_rotation_deviation_parameters=d(g["rotation_deviations"]),
                actuator_parameters_non_optimizable=d(g["actuator_non_optimizable"]),
                actuator_parameters_optimizable=d(g["actuator_optimizable"]), device=dev))
        return cls(power_plant_position=parsed["power_plant_position"].to(dev),
                   solar_tower=SolarTower([planar, cyl], device=dev), light_sources=LightSourceArray(suns),
                   heliostat_field=HeliostatField(groups, device=dev))

    def index_mapping(self, heliostat_group, string_mapping: list[tuple[str, str, torch.Tensor]] | None = None,
                      single_incident_ray_direction: torch.Tensor | None = None, single_target_area_index: int = 0,
                      device: torch.device | None = None) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """-> ``(active_heliostats_mask[Nh] int32, target_area_indices[N] int32, incident_ray_directions[N,4])``
        (``:261-418``); a heliostat named k times in ``string_mapping`` is activated k times."""
        device = get_device(device) if device is not None else heliostat_group.positions.device
        n_targets = len(self.solar_tower.target_name_to_index)
        if string_mapping is None:
            d = torch.tensor([0.0, 1.0, 0.0, 0.0]) if single_incident_ray_direction is None else \
                single_incident_ray_direction.detach().cpu().float()
            if d.shape != torch.Size([4]) or abs(float(d[3])) > 1e-8 or abs(float(torch.norm(d[:3])) - 1.0) > 1e-5:
                raise ValueError("The specified single incident ray direction is invalid. Please provide a "
                                 "normalized 4D tensor with last element 0.0.")
            if single_target_area_index >= n_targets:
                raise ValueError(f"The specified single target area index is invalid. Only {n_targets} target "
                                 "areas exist in this scenario.")
            nh = heliostat_group.number_of_heliostats
            mask = torch.ones(nh, dtype=torch.int32, device=device)
            tidx = torch.full((nh,), single_target_area_index, dtype=torch.int32, device=device)
            return mask, tidx, d.to(device).expand(nh, -1)
        rows = [m for m in string_mapping if m[0] in heliostat_group.names]
        errors = []
        for i, (_, target_name, direction) in enumerate(rows):
            if target_name not in self.solar_tower.target_name_to_index:
                errors.append(f"Invalid target '{target_name}' (Found at index {i} of provided mapping) not found "
                              "in this scenario.")
            d = direction.detach().cpu().float()
            if d.shape != torch.Size([4]) or abs(float(d[3])) > 1e-2 or abs(float(torch.norm(d)) - 1.0) > 2e-4:
                errors.append(f"Invalid incident ray direction (Found at index {i} of provided mapping). This must "
                              "be a normalized 4D tensor with last element 0.0.")
        if errors:
            raise ValueError(" ".join(errors))
        per_heliostat = defaultdict(list)
        mask_host = [0] * heliostat_group.number_of_heliostats
        name_to_row = {name: i for i, name in enumerate(heliostat_group.names)}
        for name, target_name, direction in rows:
            mask_host[name_to_row[name]] += 1
            per_heliostat[name].append((self.solar_tower.target_name_to_index[target_name], direction))
        tidx_host, dirs = [], []
        for name in heliostat_group.names:
            for t, direction in per_heliostat.get(name, []):
                tidx_host.append(t)
                dirs.append(direction.detach().cpu().float())
        mask = torch.tensor(mask_host, dtype=torch.int32, device=device)
        tidx = torch.tensor(tidx_host, dtype=torch.int32, device=device)
        inc = torch.stack(dirs).to(device) if dirs else torch.empty(0, 4, device=device)
        return mask, tidx, inc

    def set_number_of_rays(self, number_of_rays: int) -> None:
        self.light_sources.light_source_list[0].number_of_rays = number_of_rays

    def __repr__(self) -> str:
        n = sum(len(g.names) for g in self.heliostat_field.heliostat_groups)
        return (f"ARTIST-B200 Scenario: power plant at {self.power_plant_position.tolist()}, "
                f"{len(self.solar_tower.target_name_to_index)} target area(s), "
                f"{len(self.light_sources.light_source_list)} light source(s), {n} heliostat(s).")
