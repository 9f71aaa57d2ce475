"""-m gpu: the reference's exported step functions (``artist.raytracing.reflect / line_plane_intersections /
line_cylinder_intersections``, ``HeliostatRayTracer.bilinear_splatting``) as stand-alone CUDA kernels, with the reference's
signatures, against the oracle (pinned to the reference) on seeded inputs and against the reference's inline KATs
(tests/raytracing/test_geometry.py:13-83, 195-335)."""
import pytest
import torch

from oracle import artist_oracle as O
from tests import cases

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class _Planar:
    def __init__(self, tg):
        self.centers, self.normals, self.dimensions = tg.planar_centers.to(DEV), tg.planar_normals.to(DEV), tg.planar_dimensions.to(DEV)


class _Cyl:
    def __init__(self, tg):
        self.centers, self.normals, self.axes = tg.cyl_centers.to(DEV), tg.cyl_normals.to(DEV), tg.cyl_axes.to(DEV)
        self.radii, self.heights, self.opening_angles = tg.cyl_radii.to(DEV), tg.cyl_heights.to(DEV), tg.cyl_opening_angles.to(DEV)


def test_reflect_kat_and_bit_exact():
    from artist_b200.raytracing import reflect

    inc = torch.tensor([[1.0, 1.0, 1.0, 0.0], [1.0, 1.0, 1.0, 0.0], [1.0, 1.0, 1.0, 0.0], [2.0, 1.0, 3.0, 0.0]])
    nrm = torch.tensor([[0.0, 0.0, 1.0, 0.0], [0.0, 1.0, 0.0, 0.0], [1.0, 0.0, 0.0, 0.0], [0.3, 0.6, 0.7, 0.0]])
    want = torch.tensor([[1.0, 1.0, -1.0, 0.0], [1.0, -1.0, 1.0, 0.0], [-1.0, 1.0, 1.0, 0.0], [0.02, -2.96, -1.62, 0.0]])
    got = reflect(inc[:, None].to(DEV), nrm[:, None].to(DEV))
    torch.testing.assert_close(got[:, 0].cpu(), want, rtol=1e-4, atol=1e-4)
    torch.manual_seed(1)
    i, n = torch.randn(37, 1, 4), torch.randn(37, 211, 4)     # all four components, as torch sums them
    assert torch.equal(reflect(i.to(DEV), n.to(DEV)).cpu(), O.reflect(i, n))


def _rays(case):
    from artist_b200.scene import Rays

    refl = O.reflect(case["incident"][:, None], case["normals"])
    dirs = O.scatter_rays(case["dist_u"], case["dist_e"], refl)
    mags = 0.25 + torch.rand(dirs.shape[:3], generator=torch.Generator().manual_seed(3))
    return dirs, mags, Rays(dirs.to(DEV), mags.to(DEV))


@pytest.mark.parametrize("res", [(256, 256), (64, 48)])
def test_line_plane_intersections_bit_exact(res):
    from artist_b200.raytracing import line_plane_intersections

    case = cases.make_case(n=5, points_per_facet=(9, 9), rays=4, target_pattern=(0,))
    tg = case["targets"]
    dirs, mags, rays = _rays(case)
    tidx = torch.zeros(5, dtype=torch.int32)
    want = O.line_plane_intersections(dirs, mags, case["points"], tg.planar_centers, tg.planar_normals, tg.planar_dimensions,
                                      tidx, torch.tensor(res))
    got = line_plane_intersections(rays, case["points"].to(DEV), _Planar(tg), tidx.to(DEV), torch.tensor(res))
    assert (want[3] > 0).any()
    for g, w, name in zip(got, want, ("e", "u", "distance", "intensity")):
        assert torch.equal(g.cpu(), w), name
    # default target (index 0) when no indices are given; rays flying away from the plane are invalid
    again = line_plane_intersections(rays, case["points"].to(DEV), _Planar(tg), None, torch.tensor(res))
    assert all(torch.equal(a, b) for a, b in zip(again, got))
    from artist_b200.scene import Rays

    away = line_plane_intersections(Rays(-rays.ray_directions, rays.ray_magnitudes), case["points"].to(DEV), _Planar(tg), None,
                                    torch.tensor(res))
    assert (away[0] == res[0] - 1).all() and (away[1] == 0).all() and (away[2] == 0).all() and (away[3] == 0).all()


def test_line_cylinder_intersections_match_the_oracle():
    from artist_b200.raytracing import line_cylinder_intersections

    case = cases.make_case(n=4, points_per_facet=(9, 9), rays=4, target_pattern=(1,))
    tg = case["targets"]
    dirs, mags, rays = _rays(case)
    tidx = torch.zeros(4, dtype=torch.int32)                      # index within the cylindrical areas
    res = torch.tensor([128, 96])
    want = O.line_cylinder_intersections(dirs, mags, case["points"], tg.cyl_centers, tg.cyl_normals, tg.cyl_axes, tg.cyl_radii,
                                         tg.cyl_heights, tg.cyl_opening_angles, tidx, res)
    got = line_cylinder_intersections(rays, case["points"].to(DEV), _Cyl(tg), tidx.to(DEV), res)
    hit = want[3] > 0
    assert hit.any() and torch.equal(got[3].cpu() > 0, hit)
    # the reference's operation order (tests/test_gpu_trace_parity.py::test_cylindrical_and_mixed_targets): hit distance
    # and height coordinate identical up to torch's MKL sqrt, the angle up to an ulp of atan2 (round 1: 0.05 px)
    for g, w in zip(got[:2], want[:2]):
        assert (g.cpu() - w)[hit].abs().max() < 1e-3
    assert (got[2].cpu() == want[2])[hit].float().mean() >= 0.999
    assert ((got[2].cpu() - want[2])[hit].abs() / want[2][hit]).max() < 2.5e-7
    assert ((got[3].cpu() - want[3])[hit].abs() / want[3][hit]).max() < 2e-6


def test_bilinear_splatting_method_and_edge_rows():
    from artist_b200 import HeliostatRayTracer, build_synthetic_scenario

    scenario, group = build_synthetic_scenario(2, number_of_rays=3, points_per_facet=(6, 6), device=DEV)
    mask, tidx, inc = scenario.index_mapping(group)
    group.activate_heliostats(mask)
    tracer = HeliostatRayTracer(scenario, group, blocking_active=False, bitmap_resolution=torch.tensor([8, 6]))
    be = torch.tensor([[[7.0, 3.25, 0.0]]])
    bu = torch.tensor([[[2.0, 5.0, 1.5]]])
    out = tracer.bilinear_splatting(be.to(DEV), bu.to(DEV), torch.ones(1, 1, 3, device=DEV), device=DEV)
    assert out.shape == (1, 6, 8) and torch.equal(out.cpu(), O.bilinear_splatting(be, bu, torch.ones(1, 1, 3), torch.tensor([8, 6])))
    g = torch.Generator().manual_seed(5)
    be = torch.rand(3, 4, 50, generator=g) * 9 - 1          # some rays off the bitmap on every side
    bu = torch.rand(3, 4, 50, generator=g) * 7 - 1
    v = torch.rand(3, 4, 50, generator=g)
    want = O.bilinear_splatting(be, bu, v, torch.tensor([8, 6]))
    got = tracer.bilinear_splatting(be.to(DEV), bu.to(DEV), v.to(DEV), device=DEV)
    assert (got.cpu() - want).abs().max() <= 1e-6 * want.max()


def test_step_functions_compose_to_trace_rays():
    """reflect -> scatter -> line_plane_intersections -> bilinear_splatting (materialised, step by step) equals the fused
    trace_rays kernel on the same inputs: identical pixel coordinates, bitmaps within the fixed-point tolerance."""
    from artist_b200 import ops
    from artist_b200.raytracing import line_plane_intersections, reflect
    from artist_b200.raytracing.geometry import bilinear_splatting
    from artist_b200.scene import Rays
    from tests.test_gpu_trace_parity import _dev_targets

    case = cases.make_case(n=3, points_per_facet=(12, 12), rays=5, target_pattern=(0,))
    res = (96, 80)
    pts, nrm, inc = case["points"].to(DEV), case["normals"].to(DEV), case["incident"].to(DEV)
    refl = reflect(inc[:, None], nrm)
    dirs = O.scatter_rays(case["dist_u"], case["dist_e"], refl.cpu()).to(DEV)      # torch-CPU trig = strict table mode
    rays = Rays(dirs, torch.ones(dirs.shape[:3], device=DEV))
    be, bu, t, lam = line_plane_intersections(rays, pts, _Planar(case["targets"]), case["target_idx"].to(DEV), torch.tensor(res))
    stepwise = bilinear_splatting(be, bu, lam * 0.935, res)
    opt = ops.TraceOptions(res_e=res[0], res_u=res[1], trig_mode=1, scatter_sigma=2.09e-3)
    (flux, *_), (dbe, dbu, dt, dlam) = ops.trace_debug(pts, nrm, inc, ops.pack_distortions(case["dist_u"].to(DEV), case["dist_e"].to(DEV)),
                                                      case["target_idx"].to(DEV), _dev_targets(case["targets"], DEV), opt,
                                                      trig=cases.cpu_trig(case["dist_u"], case["dist_e"]).to(DEV))
    assert torch.equal(be, dbe) and torch.equal(bu, dbu) and torch.equal(t, dt) and torch.equal(lam, dlam)
    assert (flux - stepwise).abs().max() <= 1e-5 * flux.max()


def test_scatter_rays_method_matches_the_oracle():
    from artist_b200 import HeliostatRayTracer, build_synthetic_scenario

    scenario, group = build_synthetic_scenario(2, number_of_rays=3, points_per_facet=(6, 6), device=DEV)
    mask, tidx, inc = scenario.index_mapping(group)
    group.activate_heliostats(mask)
    tracer = HeliostatRayTracer(scenario, group, blocking_active=False)
    case = cases.make_case(n=3, points_per_facet=(7, 7), rays=4)
    refl = O.reflect(case["incident"][:, None], case["normals"])
    want = O.scatter_rays(case["dist_u"], case["dist_e"], refl)
    rays = tracer.scatter_rays(case["dist_u"].to(DEV), case["dist_e"].to(DEV), refl.to(DEV), device=DEV)
    assert rays.ray_directions.shape == want.shape and rays.ray_magnitudes.shape == want.shape[:3]
    assert (rays.ray_magnitudes == 1.0).all()
    # device trig is within an ulp of torch-CPU's: directions agree to 2e-7
    assert (rays.ray_directions.cpu() - want).abs().max() <= 2e-7
