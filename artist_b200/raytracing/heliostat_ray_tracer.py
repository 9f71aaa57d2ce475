from __future__ import annotations

import torch

from .. import ops
from ..scene.rays import Rays
from ..util import indices
from .sampling import DistortionsDataset, RestrictedDistributedSampler


class _BatchLoader:
    """Iterable with the DataLoader's batching of the sampler's rows (API compatibility only: the fused
    kernel reads the distortion buffer in place, nothing is collated)."""

    def __init__(self, dataset: DistortionsDataset, sampler: RestrictedDistributedSampler, batch_size: int) -> None:
        self.dataset, self.sampler, self.batch_size = dataset, sampler, batch_size

    def __iter__(self):
        rows = self.sampler.rank_indices
        for s in range(0, len(rows), self.batch_size):
            idx = torch.as_tensor(rows[s:s + self.batch_size], device=self.dataset.distortions_u.device)
            yield self.dataset.distortions_u[idx], self.dataset.distortions_e[idx]

    def __len__(self) -> int:
        return (len(self.sampler.rank_indices) + self.batch_size - 1) // self.batch_size


class HeliostatRayTracer:
    """Drop-in for ``artist.raytracing.HeliostatRayTracer`` (``heliostat_ray_tracer.py:19-778``) whose
    ``trace_rays`` is ONE fused CUDA pass (``ab200_trace_fwd``) with an explicit backward kernel.

    Same constructor, attributes, methods, shapes and side effects (the constructor reseeds the global
    torch RNG through ``Sun.get_distortions``).  Differences, all deliberate:
    rows of samples that belong to other ranks are returned as zeros instead of uninitialised memory;
    ``batch_size`` is accepted but irrelevant (no per-ray tensors exist, so memory does not grow with it);
    ``preferred_reflection_directions`` is evaluated lazily.
    """

    def __init__(self, scenario, heliostat_group, blocking_active: bool = True, world_size: int = 1, rank: int = 0,
                 batch_size: int = 100, random_seed: int = 7,
                 bitmap_resolution: torch.Tensor = torch.tensor([indices.bitmap_resolution, indices.bitmap_resolution]),
                 dni: float | None = None) -> None:
        self.scenario = scenario
        self.heliostat_group = heliostat_group
        self.blocking_active = blocking_active
        self.world_size = world_size
        self.rank = rank
        self.batch_size = batch_size
        self.light_source = scenario.light_sources.light_source_list[0]
        self.distortions_dataset = DistortionsDataset(
            light_source=self.light_source,
            number_of_points_per_heliostat=heliostat_group.active_surface_points.shape[1],
            number_of_active_heliostats=heliostat_group.number_of_active_heliostats, random_seed=random_seed)
        self.distortions_sampler = RestrictedDistributedSampler(
            number_of_samples=len(self.distortions_dataset),
            number_of_active_heliostats=int((heliostat_group.active_heliostats_mask > 0).sum()),
            world_size=world_size, rank=rank)
        self.distortions_loader = _BatchLoader(self.distortions_dataset, self.distortions_sampler, batch_size)
        self.bitmap_resolution = bitmap_resolution
        device = heliostat_group.active_surface_points.device
        self._device = device
        self._packed = ops.pack_distortions(self.distortions_dataset.distortions_u, self.distortions_dataset.distortions_e)
        n = len(self.distortions_dataset)
        rows = self.distortions_sampler.rank_indices
        self._local_rows = None if len(rows) == n else torch.tensor(rows, dtype=torch.int32, device=device)
        self._targets = ops.TargetTensors.from_solar_tower(scenario.solar_tower, device)

        if self.blocking_active:
            groups = scenario.heliostat_field.heliostat_groups
            self.blocking_heliostat_surfaces = torch.cat([g.surface_points for g in groups])
            active = []
            for g in groups:
                surfaces = g.surface_points + g.positions.unsqueeze(1)
                mask = g.active_heliostats_mask.bool()
                if mask.any():
                    surfaces[mask] = g.active_surface_points
                active.append(surfaces)
            self.blocking_heliostat_surfaces_active = torch.cat(active)

        if dni is not None:
            # heliostat area from the canting vectors of the first heliostat (:185-203)
            canting_norm = (torch.norm(heliostat_group.canting[0], dim=1)[0])[:2]
            dims = canting_norm * 4 + 0.02
            area = dims[0] * dims[1]
            rays_per_heliostat = heliostat_group.surface_points.shape[1] * self.light_source.number_of_rays
            self.ray_magnitude = dni * area / rays_per_heliostat
        else:
            self.ray_magnitude = 1.0

    # ---------------------------------------------------------------------------------------------
    def get_sampler_indices(self) -> torch.Tensor:
        return torch.tensor(self.distortions_sampler.rank_indices, device=self.distortions_dataset.distortions_u.device)

    def trace_rays(self, incident_ray_directions: torch.Tensor, active_heliostats_mask: torch.Tensor,
                   target_area_indices: torch.Tensor, ray_extinction_factor: float = 0.0,
                   mirror_reflectivity: float = 0.935, device: torch.device | None = None):
        """-> ``(flux[N,U,E], intercept[N], on_target[N], blocking[N])`` (``:220-508``); the flux is
        differentiable w.r.t. ``heliostat_group.active_surface_points / active_surface_normals``."""
        group = self.heliostat_group
        assert torch.equal(group.active_heliostats_mask, active_heliostats_mask), (
            "Some heliostats were not aligned and cannot be raytraced.")
        if self.blocking_active:
            raise NotImplementedError(
                "blocking_active=True is not built yet in artist_b200 (SURVEY.md 8f-1); pass blocking_active=False")
        group._reflection_inputs = (incident_ray_directions, group.active_surface_normals)
        sigma = getattr(self.light_source, "scatter_sigma", 0.0)
        opt = ops.TraceOptions(
            res_e=int(self.bitmap_resolution[indices.unbatched_bitmap_e]),
            res_u=int(self.bitmap_resolution[indices.unbatched_bitmap_u]),
            ray_magnitude=float(self.ray_magnitude), ray_extinction_factor=ray_extinction_factor,
            mirror_reflectivity=mirror_reflectivity, scatter_sigma=sigma)
        return ops.trace(group.active_surface_points, group.active_surface_normals, incident_ray_directions,
                         self._packed, target_area_indices, self._targets, opt, local_rows=self._local_rows)

    def get_bitmaps_per_target(self, bitmaps_per_heliostat: torch.Tensor, target_area_indices: torch.Tensor,
                               device: torch.device | None = None) -> torch.Tensor:
        """``[N,U,E]`` -> ``[T,U,E]`` (``:563-608``)."""
        n_targets = int(self.scenario.solar_tower.number_of_target_areas_per_type.sum())
        return ops.bitmaps_per_target(bitmaps_per_heliostat, target_area_indices, n_targets)

    # The helper below exists for API parity; trace_rays does not go through it.
    def scatter_rays(self, distortion_u, distortion_e, original_ray_direction, device=None) -> Rays:
        """Scattered directions ``[B,R,P,4]`` around the preferred directions (``:510-561``)."""
        ce, se, cu, su = torch.cos(distortion_e), torch.sin(distortion_e), torch.cos(distortion_u), torch.sin(distortion_u)
        r = original_ray_direction.unsqueeze(1)
        dx = cu * r[..., 0] + (-su) * r[..., 1]
        dy = (ce * su) * r[..., 0] + (ce * cu) * r[..., 1] + (-se) * r[..., 2]
        dz = (se * su) * r[..., 0] + (se * cu) * r[..., 1] + ce * r[..., 2]
        dirs = torch.stack([dx, dy, dz, r[..., 3].expand_as(dx)], dim=-1)
        return Rays(dirs, torch.full(dirs.shape[:3], float(self.ray_magnitude), device=dirs.device))
