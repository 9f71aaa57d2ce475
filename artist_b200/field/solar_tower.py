from __future__ import annotations

from collections.abc import Sequence

import torch

from ..util.env import get_device
from .tower_target_areas import TowerTargetAreas


class SolarTower:
    """Tower with its target areas; global index = planar areas first, then cylindrical
    (``artist/field/solar_tower.py:19-188``)."""

    def __init__(self, target_areas: Sequence[TowerTargetAreas], device: torch.device | None = None) -> None:
        device = get_device(device)
        self.target_areas = target_areas
        self.number_of_target_area_types = len(target_areas)
        self.number_of_target_areas_per_type = torch.tensor([t.number_of_target_areas for t in target_areas],
                                                            device=device)
        self.target_name_to_index = {}
        self.index_to_target_area = []
        for area_type in target_areas:
            for local, name in enumerate(area_type.names):
                self.target_name_to_index[name] = len(self.target_name_to_index)
                self.index_to_target_area.append((area_type, local))
        self._n_planar = target_areas[0].number_of_target_areas if target_areas else 0

    def get_centers_of_target_areas(self, target_area_indices: torch.Tensor,
                                    device: torch.device | None = None) -> torch.Tensor:
        """Aim points ``[N,4]``: plane centre, or cylinder centre + radius * normal (``:129-188``)."""
        planar, cyl = self.target_areas[0], self.target_areas[1]
        idx = target_area_indices.long()
        centers = torch.zeros(idx.shape[0], 4, device=target_area_indices.device)
        is_planar = idx < self._n_planar
        if planar.number_of_target_areas > 0:
            pi = idx.clamp(max=max(planar.number_of_target_areas - 1, 0))
            centers = torch.where(is_planar[:, None], planar.centers.to(centers.device)[pi], centers)
        if cyl.number_of_target_areas > 0:
            ci = (idx - self._n_planar).clamp(min=0)
            cc = cyl.centers.to(centers.device)[ci] + cyl.radii.to(centers.device)[ci][:, None] * cyl.normals.to(centers.device)[ci]
            centers = torch.where(is_planar[:, None], centers, cc)
        centers[:, 3] = 1.0
        return centers
