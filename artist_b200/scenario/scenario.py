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
