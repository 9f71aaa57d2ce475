"""Flux-bitmap utilities that follow the ray tracer in the reference's optimisers (``artist/flux/bitmap.py``), same
signatures; the per-pixel work runs in ``csrc/flux.cu`` with hand-written backward kernels."""
from __future__ import annotations

import torch

from .. import ops
from ..util.env import get_device

UTIS_CROP_WIDTH = 6      # artist/util/constants.py:217-219
UTIS_CROP_HEIGHT = 6


def get_center_of_mass(bitmaps: torch.Tensor, device: torch.device | None = None) -> torch.Tensor:
    """Centre of mass ``[N,2]`` = (e, u) in pixel units of ``[N,U,E]`` bitmaps (``bitmap.py:12-55``)."""
    return ops.flux_center_of_mass(bitmaps, normalised=False)


def trapezoid_distribution(total_width: int, slope_width: int, plateau_width: int,
                           device: torch.device | None = None) -> torch.Tensor:
    """1-D trapezoid window (``bitmap.py:58-118``): a handful of values, built with tensor ops on ``device``."""
    device = get_device(device)
    distances = torch.abs(torch.arange(total_width, device=device) - (total_width - 1) / 2.0) - plateau_width / 2.0
    if slope_width == 0:
        return (distances <= 0).to(dtype=torch.float32)
    return 1 - (distances / slope_width).clamp(min=0, max=1)


def crop_flux_distributions_around_center(flux_distributions: torch.Tensor, solar_tower, target_area_indices: torch.Tensor,
                                          crop_width: float = UTIS_CROP_WIDTH, crop_height: float = UTIS_CROP_HEIGHT,
                                          device: torch.device | None = None) -> torch.Tensor:
    """Crop ``[N,U,E]`` flux bitmaps to ``crop_width x crop_height`` metres around their centres of mass, resampled to the
    same resolution (``bitmap.py:121-246``).  Target extents: planar areas their plane dimensions, cylindrical areas
    radius x opening angle by height."""
    dev = flux_distributions.device
    idx = target_area_indices.to(dev).long()
    planar, cyl = solar_tower.target_areas[0], solar_tower.target_areas[1]
    n_planar = planar.number_of_target_areas
    dims = torch.empty(idx.shape[0], 2, device=dev)
    is_planar = idx < n_planar
    if n_planar > 0:
        dims = torch.where(is_planar[:, None], planar.dimensions.to(dev).float()[idx.clamp(max=n_planar - 1)], dims)
    if cyl.number_of_target_areas > 0:
        ci = (idx - n_planar).clamp(min=0)
        cdims = torch.stack([cyl.radii.to(dev).float()[ci] * cyl.opening_angles.to(dev).float()[ci],
                             cyl.heights.to(dev).float()[ci]], dim=1)
        dims = torch.where(is_planar[:, None], dims, cdims)
    scale = torch.stack([crop_width / dims[:, 0].clamp(min=1e-8), crop_height / dims[:, 1].clamp(min=1e-8)], dim=1)
    return ops.flux_crop_around_center(flux_distributions, scale)
