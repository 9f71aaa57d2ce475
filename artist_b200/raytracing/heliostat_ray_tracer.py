from __future__ import annotations

import math

import torch

from .. import _lib, ops
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
            number_of_points_per_heliostat=heliostat_group._active_points_per_heliostat(),
            number_of_active_heliostats=heliostat_group.number_of_active_heliostats, random_seed=random_seed)
        self.distortions_sampler = RestrictedDistributedSampler(
            number_of_samples=len(self.distortions_dataset),
            number_of_active_heliostats=int((heliostat_group.active_heliostats_mask > 0).sum()),
            world_size=world_size, rank=rank)
        self.distortions_loader = _BatchLoader(self.distortions_dataset, self.distortions_sampler, batch_size)
        self.bitmap_resolution = bitmap_resolution
        device = heliostat_group.surface_points.device
        self._device = device
        self._packed = ops.pack_distortions(self.distortions_dataset.distortions_u, self.distortions_dataset.distortions_e)
        n = len(self.distortions_dataset)
        rows = self.distortions_sampler.rank_indices
        self._local_rows = None if len(rows) == n else torch.tensor(rows, dtype=torch.int32, device=device)
        self._targets = ops.TargetTensors.from_solar_tower(scenario.solar_tower, device)

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
        group._reflection_inputs = incident_ray_directions
        sigma = getattr(self.light_source, "scatter_sigma", 0.0)
        opt = ops.TraceOptions(
            res_e=int(self.bitmap_resolution[indices.unbatched_bitmap_e]),
            res_u=int(self.bitmap_resolution[indices.unbatched_bitmap_u]),
            ray_magnitude=float(self.ray_magnitude), ray_extinction_factor=ray_extinction_factor,
            mirror_reflectivity=mirror_reflectivity, scatter_sigma=sigma,
            one_cta_per_sample=getattr(self, "_force_one_cta_per_sample", False))
        blocking = self._blocking_inputs(target_area_indices) if self.blocking_active else None
        if torch.is_grad_enabled() and self._targets.n_planar + self._targets.n_cyl > 1:
            ops.prefetch_uniform_target(target_area_indices)   # for get_bitmaps_per_target's backward, off the critical path
        fused = group._fused_alignment(with_map=True)
        if fused is not None:   # alignment not materialised: the kernels rotate the un-aligned rows themselves - and, with
            points, normals, orientations, amap = fused    # an activation map, read the group's un-replicated rows
        elif group._pending_gather is not None and group._asp is None and group._asn is None:
            points, normals, orientations, amap = group.surface_points, group.surface_normals, None, group._pending_gather
        else:
            points, normals, orientations, amap = group.active_surface_points, group.active_surface_normals, None, None
        return ops.trace(points, normals, incident_ray_directions, self._packed, target_area_indices, self._targets,
                         opt, local_rows=self._local_rows, blocking=blocking, orientations=orientations, activation=amap)

    # ---- blocking (artist/raytracing/blocking.py, heliostat_ray_tracer.py:159-183,292-301,445-480) -----------------
    @staticmethod
    def _corner_rows(number_of_points: int) -> list[int]:
        """The four fixed surface-point indices the reference uses as rectangle corners (``blocking.py:176-193``;
        assumes four equal square facets): lower left, upper left, upper right, lower right."""
        q = int(math.sqrt(number_of_points / 4))
        return [number_of_points // 2, q - 1, number_of_points // 2 - 1, number_of_points - q]

    def _blocker_corners(self) -> tuple[torch.Tensor, torch.Tensor]:
        """Corner points ``[H_all,4,4]`` of every heliostat of every group (aligned where a group is active,
        horizontal at its position otherwise - ``:170-183``) and, for this tracer's group, the primitive row of each
        active sample.  Only the 4 corner points per heliostat are gathered; the reference concatenates all surfaces."""
        corners, owner, offset = [], None, 0
        for g in self.scenario.heliostat_field.heliostat_groups:
            # (no host->device copy and no device->host read per call: the corner-row indices are cached, "is any heliostat of
            # the group active" is the count activate_heliostats already holds on the host - keeps trace_rays capturable
            # in a CUDA graph and off the launch pipeline's critical path)
            key = (g.surface_points.shape[1], g.surface_points.device)
            cache = self.__dict__.setdefault("_corner_rows_cache", {})
            rows = cache.get(key)
            if rows is None:
                rows = cache[key] = torch.tensor(self._corner_rows(key[0]), device=key[1])
            c = g.surface_points.index_select(1, rows) + g.positions.unsqueeze(1)
            if g.number_of_active_heliostats > 0:
                active_rows = getattr(g, "_active_rows", None)
                if active_rows is None:     # all-ones mask: sample i is heliostat i
                    active_rows = torch.arange(g.number_of_heliostats, device=c.device)
                fused = g._fused_alignment(with_map=True)
                if fused is not None:   # only the 4 corner rows are rotated; the full alignment stays lazy
                    cp, cn = fused[0].index_select(1, rows), fused[1].index_select(1, rows)
                    if fused[3] is not None:
                        cp, cn = cp.index_select(0, fused[3].rows.long()), cn.index_select(0, fused[3].rows.long())
                    aligned, _ = ops.align_surfaces(cp, cn, fused[2])
                else:
                    aligned = g.active_surface_points.index_select(1, rows)
                # a heliostat activated several times contributes the geometry of its last sample
                c = c.index_copy(0, active_rows.long(), aligned)
                if g is self.heliostat_group:
                    owner = (active_rows + offset).to(torch.int32)
            corners.append(c)
            offset += g.number_of_heliostats
        return torch.cat(corners), owner

    def _blocking_inputs(self, target_area_indices: torch.Tensor) -> ops.BlockingInputs:
        corners, owner = self._blocker_corners()
        spans = torch.stack([corners[:, 1] - corners[:, 0], corners[:, 3] - corners[:, 0]], dim=1)
        normals = torch.nn.functional.normalize(torch.linalg.cross(spans[:, 0, :3], spans[:, 1, :3], dim=-1), dim=-1)
        tower = self.scenario.solar_tower
        aim = tower.get_centers_of_target_areas(target_area_indices)
        planar, cyl = tower.target_areas[0], tower.target_areas[1]
        n_planar = planar.number_of_target_areas
        idx = target_area_indices.long()
        radius = torch.zeros(idx.shape[0], device=aim.device)
        if n_planar > 0:
            pr = 0.5 * torch.linalg.norm(planar.dimensions.to(aim.device).float(), dim=1)
            radius = torch.where(idx < n_planar, pr[idx.clamp(max=n_planar - 1)], radius)
        if cyl.number_of_target_areas > 0:
            cr = 1.5 * torch.maximum(cyl.radii.to(aim.device).float(), 0.5 * cyl.heights.to(aim.device).float())
            radius = torch.where(idx < n_planar, radius, cr[(idx - n_planar).clamp(min=0)])
        return ops.BlockingInputs(corners=corners, spans=spans, normals=normals, sample_to_blocker=owner, aim_points=aim,
                                  target_radius=radius)

    @property
    def blocking_heliostat_surfaces(self) -> torch.Tensor:
        return torch.cat([g.surface_points for g in self.scenario.heliostat_field.heliostat_groups])

    @property
    def blocking_heliostat_surfaces_active(self) -> torch.Tensor:
        out = []
        for g in self.scenario.heliostat_field.heliostat_groups:
            surfaces = g.surface_points + g.positions.unsqueeze(1)
            mask = g.active_heliostats_mask.bool()
            if mask.any():
                surfaces = surfaces.clone()
                surfaces[mask] = g.active_surface_points
            out.append(surfaces)
        return torch.cat(out)

    def get_bitmaps_per_target(self, bitmaps_per_heliostat: torch.Tensor, target_area_indices: torch.Tensor,
                               device: torch.device | None = None) -> torch.Tensor:
        """``[N,U,E]`` -> ``[T,U,E]`` (``:563-608``)."""
        # the count lives in a device tensor (as upstream): read it once per tensor version, not once per call (a
        # device->host read in the middle of every step stalls the launch pipeline)
        per_type = self.scenario.solar_tower.number_of_target_areas_per_type
        key = (id(per_type), per_type._version)
        if getattr(self, "_n_targets_key", None) != key:
            self._n_targets, self._n_targets_key, self._n_targets_ref = int(per_type.sum()), key, per_type
        return ops.bitmaps_per_target(bitmaps_per_heliostat, target_area_indices, self._n_targets)

    def bilinear_splatting(self, bitmap_intersections_e: torch.Tensor, bitmap_intersections_u: torch.Tensor,
                           absolute_intensities: torch.Tensor, device: torch.device | None = None) -> torch.Tensor:
        """Stand-alone splat of materialised per-ray coordinates (``:610-778``); ``trace_rays`` splats inside its
        fused kernel instead."""
        from .geometry import bilinear_splatting

        return bilinear_splatting(bitmap_intersections_e, bitmap_intersections_u, absolute_intensities,
                                  self.bitmap_resolution)

    # The helper below exists for API parity; trace_rays does not go through it.
    def scatter_rays(self, distortion_u, distortion_e, original_ray_direction, device=None) -> Rays:
        """Scattered directions ``[B,R,P,4]`` around the preferred directions (``:510-561``) - ``ab200_scatter_rays``."""
        du = ops._f32(distortion_u.detach(), "distortion_u")
        de = ops._f32(distortion_e.detach(), "distortion_e")
        refl = ops._f32(original_ray_direction.detach(), "original_ray_direction")
        b, r, p = du.shape
        dirs = torch.empty(b, r, p, 4, device=du.device)
        _lib.call("ab200_scatter_rays", ops._p(du), ops._p(de), ops._p(refl), b, r, p, ops._p(dirs), ops._stream())
        return Rays(dirs, torch.full((b, r, p), float(self.ray_magnitude), device=dirs.device))
