"""-m gpu: flux-bitmap centre of mass and centre-of-mass crop (artist/flux/bitmap.py) - the CUDA kernels
through the reference-shaped functions against the REAL reference's outputs (tests/golden/flux_golden.pt) and the
oracle's autograd."""
import os

import pytest
import torch

from oracle import artist_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "flux_golden.pt")


class _Areas:
    def __init__(self, **kw):
        self.__dict__.update(kw)


def _tower():
    planar = _Areas(number_of_target_areas=2, dimensions=torch.tensor([[8.0, 7.0], [5.4, 6.4]], device=DEV))
    cyl = _Areas(number_of_target_areas=1, radii=torch.tensor([4.14], device=DEV), heights=torch.tensor([5.229], device=DEV),
                 opening_angles=torch.tensor([1.0472], device=DEV))
    return _Areas(target_areas=[planar, cyl])


@pytest.mark.parametrize("key", ["small", "square"])
def test_center_of_mass_and_crop_against_the_reference(key):
    from artist_b200.flux import crop_flux_distributions_around_center, get_center_of_mass

    c = torch.load(GOLDEN, weights_only=False)[key]
    flux = c["flux"].to(DEV)
    com = get_center_of_mass(flux, device=DEV)
    assert com.shape == c["center_of_mass"].shape
    assert (com.cpu() - c["center_of_mass"]).abs().max() <= 2e-5          # pixels; fp32 sums over the bitmap
    crop = crop_flux_distributions_around_center(flux, _tower(), c["target_idx"].to(DEV), crop_width=c["crop"][0],
                                                 crop_height=c["crop"][1], device=DEV)
    assert crop.shape == c["cropped"].shape
    assert (crop.cpu() - c["cropped"]).abs().max() <= 2e-5 * c["cropped"].max()


def test_crop_and_center_of_mass_gradients_match_autograd():
    from artist_b200 import ops
    from artist_b200.flux import get_center_of_mass

    c = torch.load(GOLDEN, weights_only=False)["small"]
    torch.manual_seed(0)
    wgt = torch.rand_like(c["flux"])
    ref_in = c["flux"].clone().requires_grad_(True)
    (O.crop_flux_around_center(ref_in, c["target_dimensions"], 6, 6) * wgt).sum().backward()
    got_in = c["flux"].to(DEV).requires_grad_(True)
    scale = torch.stack([6 / c["target_dimensions"][:, 0], 6 / c["target_dimensions"][:, 1]], dim=1).to(DEV)
    (ops.flux_crop_around_center(got_in, scale) * wgt.to(DEV)).sum().backward()
    scale_ref = ref_in.grad.abs().max()
    assert (got_in.grad.cpu() - ref_in.grad).abs().max() <= 2e-4 * scale_ref
    # centre of mass
    ref_in = c["flux"].clone().requires_grad_(True)
    w2 = torch.tensor([[0.3, -1.1]])
    (O.flux_center_of_mass(ref_in) * w2).sum().backward()
    got_in = c["flux"].to(DEV).requires_grad_(True)
    (get_center_of_mass(got_in) * w2.to(DEV)).sum().backward()
    assert (got_in.grad.cpu() - ref_in.grad).abs().max() <= 1e-5 * ref_in.grad.abs().max()


def test_crop_of_traced_flux_is_differentiable_to_the_surface():
    """trace_rays -> crop -> loss -> backward reaches the surface points (the surface reconstructor's chain)."""
    from artist_b200 import HeliostatRayTracer, build_synthetic_scenario
    from artist_b200.flux import crop_flux_distributions_around_center

    scenario, group = build_synthetic_scenario(3, number_of_rays=4, points_per_facet=(10, 10), device=DEV)
    mask, tidx, inc = scenario.index_mapping(group)
    group.activate_heliostats(mask)
    pts = group.active_surface_points.clone().requires_grad_(True)
    group.active_surface_points = pts
    group.align_surfaces_with_incident_ray_directions(scenario.solar_tower.get_centers_of_target_areas(tidx), inc, mask)
    tracer = HeliostatRayTracer(scenario, group, blocking_active=False, bitmap_resolution=torch.tensor([64, 64]))
    flux, *_ = tracer.trace_rays(inc, mask, tidx)
    crop = crop_flux_distributions_around_center(flux, scenario.solar_tower, tidx, crop_width=4.0, crop_height=4.0)
    assert crop.shape == flux.shape and crop.sum() > 0
    (crop * torch.linspace(0, 1, 64, device=DEV)).sum().backward()
    assert torch.isfinite(pts.grad).all() and pts.grad.abs().max() > 0
