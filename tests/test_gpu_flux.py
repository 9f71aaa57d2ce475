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


LOSS_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loss_golden.pt")


@pytest.mark.parametrize("key", ["small", "square"])
@pytest.mark.parametrize("name", ["pixel", "kl"])
def test_bitmap_losses_against_the_reference(key, name):
    """PixelLoss / KLDivergenceLoss (artist/optim/loss.py:251-410) as fused kernels against the REAL reference's values and
    autograd gradients (tests/golden/loss_golden.pt): 2e-6 relative on the loss, 1e-5 of the largest gradient entry."""
    from artist_b200.optim import KLDivergenceLoss, PixelLoss

    c = torch.load(LOSS_GOLDEN, weights_only=False)[key]
    fn = PixelLoss() if name == "pixel" else KLDivergenceLoss()
    pred = c["prediction"].to(DEV).requires_grad_(True)
    loss = fn(prediction=pred, ground_truth=c["ground_truth"].to(DEV), reduction_dimensions=(1, 2))
    assert loss.shape == c[name].shape
    torch.testing.assert_close(loss.detach().cpu(), c[name], atol=1e-6, rtol=2e-6)
    (loss * c["weights"].to(DEV)).sum().backward()
    want = c[name + "_grad"]
    assert (pred.grad.cpu() - want).abs().max() <= 1e-5 * want.abs().max()


def test_bitmap_loss_kats_and_keyword_contract():
    """The reference's own known-answer tests (tests/optim/test_loss_functions.py:266-445) and its error messages."""
    from artist_b200.optim import KLDivergenceLoss, PixelLoss
    from tests.test_oracle_kat import KL_KATS, PIXEL_KATS

    for pred, gt, want in PIXEL_KATS:
        got = PixelLoss()(prediction=torch.tensor(pred, device=DEV), ground_truth=torch.tensor(gt, device=DEV),
                          reduction_dimensions=(1, 2))
        torch.testing.assert_close(got.cpu(), torch.tensor(want), atol=1e-6, rtol=1e-6)
    for pred, gt, want in KL_KATS:
        got = KLDivergenceLoss()(prediction=torch.tensor(pred, device=DEV), ground_truth=torch.tensor(gt, device=DEV),
                                 target_area_indices=torch.tensor([0, 1], device=DEV), reduction_dimensions=(1, 2), device=DEV)
        torch.testing.assert_close(got.cpu(), torch.tensor(want), atol=1e-6, rtol=1e-6)
    x = torch.ones(1, 2, 2, device=DEV)
    with pytest.raises(ValueError, match="The vector loss expects"):
        PixelLoss()(prediction=x, ground_truth=x)
    with pytest.raises(ValueError, match="The KL-divergence loss expects 'reduction_dimensions'"):
        KLDivergenceLoss()(prediction=x, ground_truth=x)
    with pytest.raises(ValueError, match="whole bitmap"):
        PixelLoss()(prediction=x, ground_truth=x, reduction_dimensions=(1,))


def test_bitmap_loss_at_bench_size_matches_the_oracle():
    """256 x 256 bitmaps, 64 samples: fp32 sums over 65536 pixels stay within 5e-6 of the oracle (float64 accumulate)."""
    from artist_b200 import ops

    torch.manual_seed(5)
    gt = torch.rand(64, 256, 256) ** 4
    pred = (gt * (0.5 + torch.rand(64, 256, 256))).requires_grad_(True)
    for kind, fn in ((ops.LOSS_PIXEL, O.pixel_loss), (ops.LOSS_KL_DIVERGENCE, O.kl_divergence_loss)):
        want = fn(pred.detach().double(), gt.double())
        p = pred.detach().to(DEV).requires_grad_(True)
        got = ops.flux_loss(p, gt.to(DEV), kind)
        torch.testing.assert_close(got.cpu().double(), want, atol=1e-7, rtol=5e-6)
        got.sum().backward()
        pd = pred.detach().double().requires_grad_(True)
        fn(pd, gt.double()).sum().backward()
        assert (p.grad.cpu().double() - pd.grad).abs().max() <= 2e-5 * pd.grad.abs().max()


@pytest.mark.parametrize("res", [(256, 256), (48, 64), (50, 37)])
def test_column_marching_crop_kernels_equal_the_per_pixel_kernels(res, monkeypatch):
    """Round-2 crop kernels (one thread walks a column, source rows kept in registers) against the per-pixel kernels of
    round 1 (``AB200_FLUX_PER_PIXEL``), which the tests above pin to the real reference: forward bit-identical, gradients
    equal up to summation order - for crops smaller and larger than the target, strong magnification (the per-bitmap
    fallback), and centres of mass near the bitmap's corners."""
    from artist_b200 import ops

    u, e = res
    gen = torch.Generator().manual_seed(9)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, u), torch.linspace(-1, 1, e), indexing="ij")
    centres = [(0.0, 0.0), (0.6, -0.5), (-0.85, 0.8), (0.1, 0.95), (-0.3, 0.2), (0.9, 0.9)]
    flux = torch.stack([torch.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / 0.05) + 0.01 * torch.rand(u, e, generator=gen)
                        for cx, cy in centres]).to(DEV)
    scale = torch.tensor([[0.75, 0.75], [1.3, 0.6], [0.45, 1.7], [0.3, 0.3], [1.0, 1.0], [2.5, 0.9]], device=DEV)
    wgt = torch.rand(flux.shape, generator=gen).to(DEV)

    def run():
        x = flux.clone().requires_grad_(True)
        out = ops.flux_crop_around_center(x, scale)
        (out * wgt).sum().backward()
        return out.detach(), x.grad

    new_out, new_grad = run()
    monkeypatch.setenv("AB200_FLUX_PER_PIXEL", "1")
    old_out, old_grad = run()
    assert old_out.abs().max() > 0
    assert torch.equal(new_out, old_out)
    assert (new_grad - old_grad).abs().max() <= 2e-6 * old_grad.abs().max()
