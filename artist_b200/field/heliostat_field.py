from __future__ import annotations

import math
from collections.abc import Sequence

import torch

from ..nurbs import NURBSSurfaces, create_nurbs_evaluation_grid
from .heliostat_group import HeliostatGroup


class HeliostatField:
    """All heliostat groups of a scenario (``artist/field/heliostat_field.py:24-503``)."""

    def __init__(self, heliostat_groups: Sequence[HeliostatGroup], device: torch.device | None = None) -> None:
        self.heliostat_groups = heliostat_groups
        self.number_of_heliostat_groups = len(heliostat_groups)
        self.number_of_heliostats_per_group = torch.tensor([g.number_of_heliostats for g in heliostat_groups],
                                                           device=device)

    def update_surfaces(self, device: torch.device | None = None) -> None:
        """Re-evaluate every group's surface from its (detached) control points (``:437-503``)."""
        for group in self.heliostat_groups:
            dev = group.surface_points.device
            per_facet = int(math.sqrt(group.surface_points.shape[1] / group.number_of_facets_per_heliostat))
            grid = create_nurbs_evaluation_grid(torch.tensor([per_facet, per_facet]), device=dev)
            ev = grid[None, None].expand(group.number_of_heliostats, group.number_of_facets_per_heliostat, -1, -1)
            surf = NURBSSurfaces(degrees=group.nurbs_degrees, control_points=group.nurbs_control_points.detach(), device=dev)
            pts, nrm = surf.calculate_surface_points_and_normals(ev, group.canting, group.facet_translations)
            n = group.surface_points.shape[0]
            group.surface_points = pts.reshape(n, -1, 4).detach()
            group.surface_normals = nrm.reshape(n, -1, 4).detach()
