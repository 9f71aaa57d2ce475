"""-m gpu: blocking (soft ray/rectangle occlusion) fused into the trace kernels vs the oracle, which reproduces the
reference's per-batch LBVH filter + soft mask bit for bit (tests/test_oracle_golden.py::test_blocking_*)."""
import pytest
import torch

from oracle import artist_oracle as O
from tests import cases

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _scene(n=9, ppf=(10, 10), rays=4, seed=0):
    """3x3 heliostats at 3.4 m pitch aiming at a target 6 m above ground: the back rows are heavily shadowed."""
    from artist_b200 import build_synthetic_scenario, synthetic_field_tensors

    ft = synthetic_field_tensors(n, control_points=(6, 6), surface_bump=0.001, pitch=3.4, planar_center=(0.0, 0.0, 6.0), seed=seed)
    scenario, group = build_synthetic_scenario(n, number_of_rays=rays, points_per_facet=ppf, device=DEV, field_tensors=ft)
    mask, tidx, inc = scenario.index_mapping(group, single_incident_ray_direction=torch.tensor([0.0, 0.8, -0.6, 0.0]))
    group.activate_heliostats(mask)
    group.align_surfaces_with_incident_ray_directions(scenario.solar_tower.get_centers_of_target_areas(tidx), inc, mask)
    return ft, scenario, group, mask, tidx, inc


def _oracle_blocking(points_cpu, n):
    c, s, nn = O.blocking_primitives(points_cpu)
    return dict(corners=c, spans=s, normals=nn, sample_to_blocker=torch.arange(n))


@pytest.mark.parametrize("res", [(48, 48), (128, 96)])
def test_blocked_flux_and_factors_match_oracle(res):
    from artist_b200 import HeliostatRayTracer, ops

    ft, scenario, group, mask, tidx, inc = _scene()
    tracer = HeliostatRayTracer(scenario, group, blocking_active=True, bitmap_resolution=torch.tensor(res))
    flux, ic, ot, bl = tracer.trace_rays(inc, mask, tidx)
    assert int(ops.last_blocking_overflow.item()) == 0
    pts, nrm = group.active_surface_points.cpu(), group.active_surface_normals.cpu()
    tg = cases.targets_from(ft)
    du, de = tracer.distortions_dataset.distortions_u.cpu(), tracer.distortions_dataset.distortions_e.cpu()
    ref, ric, rot, rbl = O.trace_rays(pts, nrm, inc.cpu(), du, de, tidx.cpu(), tg, res, blocking=_oracle_blocking(pts, 9))
    unblocked, *_ = O.trace_rays(pts, nrm, inc.cpu(), du, de, tidx.cpu(), tg, res)
    assert rbl.min() < 0.5 and (unblocked.sum() - ref.sum()) > 0.2 * unblocked.sum(), "scene must be strongly shadowed"
    # the evaluated rays follow the reference's operation order (block_tuv_strict): no ray changes side of the 1e-3
    # threshold, the factors are exact; measured flux error 2e-6 .. 4e-6 of peak (round 1: 2e-4 and 2.5e-3)
    ferr = (flux.cpu() - ref).abs().max() / ref.max()
    assert ferr <= 2e-5, f"flux {ferr:.3e}"
    assert (bl.cpu() - rbl).abs().max() <= 1e-6, f"blocking factor {(bl.cpu() - rbl).abs().max():.3e}"
    assert (ic.cpu() - ric).abs().max() <= 1e-6 and (ot.cpu() - rot).abs().max() <= 1e-6
    # blocking switched off gives the unshadowed flux
    plain = HeliostatRayTracer(scenario, group, blocking_active=False, bitmap_resolution=torch.tensor(res))
    f0, *_ = plain.trace_rays(inc, mask, tidx)
    assert (f0.cpu() - unblocked).abs().max() <= 1e-4 * unblocked.max()


def test_blocking_gradients_match_oracle_autograd():
    """Gradients w.r.t. the traced surfaces (ray origins / directions through the soft mask) and w.r.t. the blockers'
    geometry (their corner points are rows of the same aligned surfaces) vs autograd of the oracle."""
    from artist_b200 import HeliostatRayTracer

    res = (48, 48)
    ft, scenario, group, mask, tidx, inc = _scene(rays=6)
    pts_leaf = group.active_surface_points.detach().clone().requires_grad_(True)
    nrm_leaf = group.active_surface_normals.detach().clone().requires_grad_(True)
    group.active_surface_points, group.active_surface_normals = pts_leaf, nrm_leaf
    tracer = HeliostatRayTracer(scenario, group, blocking_active=True, bitmap_resolution=torch.tensor(res))
    flux, *_ = tracer.trace_rays(inc, mask, tidx)
    torch.manual_seed(2)
    wgt = torch.rand(9, res[1], res[0])
    (flux * wgt.to(DEV)).sum().backward()
    p = pts_leaf.detach().cpu().requires_grad_(True)
    n = nrm_leaf.detach().cpu().requires_grad_(True)
    c, s, nn = O.blocking_primitives(p)
    blk = dict(corners=c, spans=s, normals=nn, sample_to_blocker=torch.arange(9))
    tg = cases.targets_from(ft)
    ref, *_ = O.trace_rays(p, n, inc.cpu(), tracer.distortions_dataset.distortions_u.cpu(),
                           tracer.distortions_dataset.distortions_e.cpu(), tidx.cpu(), tg, res, blocking=blk)
    (ref * wgt).sum().backward()
    for got, want, name in ((pts_leaf.grad.cpu(), p.grad, "points"), (nrm_leaf.grad.cpu(), n.grad, "normals")):
        scale = want.abs().max()
        err = (got - want).abs().max() / scale
        assert err <= 5e-5, f"grad {name}: {err:.3e}"     # measured 1.5e-6 / 6.6e-6 (round 1: 2e-3)
    # the corner rows carry the blocker-geometry gradient: compare them separately (they are tiny next to the rest)
    rows = torch.tensor(HeliostatRayTracer._corner_rows(p.shape[1]))
    gc, wc = pts_leaf.grad.cpu()[:, rows], p.grad[:, rows]
    assert (gc - wc).abs().max() <= 5e-5 * wc.abs().max()     # measured 1.5e-6 (round 1: 5e-3)


def test_distant_heliostats_do_not_block():
    """The benchmark-style field (5 m pitch, 50 m high target) has no shadowing: blocking on == blocking off."""
    from artist_b200 import HeliostatRayTracer, build_synthetic_scenario

    scenario, group = build_synthetic_scenario(16, number_of_rays=4, points_per_facet=(10, 10), device=DEV)
    mask, tidx, inc = scenario.index_mapping(group)
    group.activate_heliostats(mask)
    group.align_surfaces_with_incident_ray_directions(scenario.solar_tower.get_centers_of_target_areas(tidx), inc, mask)
    a = HeliostatRayTracer(scenario, group, blocking_active=True, bitmap_resolution=torch.tensor([64, 64])).trace_rays(inc, mask, tidx)
    b = HeliostatRayTracer(scenario, group, blocking_active=False, bitmap_resolution=torch.tensor([64, 64])).trace_rays(inc, mask, tidx)
    assert torch.equal(a[0], b[0]) and torch.equal(a[3], torch.ones_like(a[3]))


def test_blocker_gradients_are_reproducible_run_to_run():
    """The gradient w.r.t. the blockers' geometry is summed per CTA in shared-memory double accumulators and across the
    samples in a fixed order (``blocker_grad_reduce_kernel``), not with global float atomics: two runs give the same bits
    (as do the bitmaps and every other gradient)."""
    from artist_b200 import HeliostatRayTracer

    res = (64, 64)
    ft, scenario, group, mask, tidx, inc = _scene(n=16, ppf=(16, 16), rays=8)
    base_p, base_n = group.active_surface_points.detach().clone(), group.active_surface_normals.detach().clone()
    torch.manual_seed(3)
    wgt = torch.rand(16, res[1], res[0], device=DEV)
    runs = []
    for _ in range(3):
        p, n = base_p.clone().requires_grad_(True), base_n.clone().requires_grad_(True)
        group.active_surface_points, group.active_surface_normals = p, n
        tracer = HeliostatRayTracer(scenario, group, blocking_active=True, bitmap_resolution=torch.tensor(res))
        flux, ic, ot, bl = tracer.trace_rays(inc, mask, tidx)
        (flux * wgt).sum().backward()
        runs.append((flux.detach().clone(), p.grad.clone(), n.grad.clone()))
    assert float(bl.min()) < 0.95, "the scene was meant to be shadowed"
    rows = torch.tensor(HeliostatRayTracer._corner_rows(base_p.shape[1]), device=DEV)
    assert runs[0][1][:, rows].abs().max() > 0
    for other in runs[1:]:
        for a, b in zip(runs[0], other):
            assert torch.equal(a, b)


def test_blocking_one_cta_per_sample_kernels_equal_split_mode():
    """The kernels a full field runs (one CTA per sample: both blocking passes, the packed second pass, per-CTA blocker-gradient
    rows) against the several-CTAs-per-sample kernels a small scene picks: identical bitmaps and factors (integer
    accumulation), gradients equal up to summation order - incl. the blockers' corner rows - and both within the usual bars
    of the oracle (asserted for the split mode by the tests above)."""
    from artist_b200 import HeliostatRayTracer

    res = (96, 80)
    ft, scenario, group, mask, tidx, inc = _scene(n=16, ppf=(20, 20), rays=6)
    base_p, base_n = group.active_surface_points.detach().clone(), group.active_surface_normals.detach().clone()
    torch.manual_seed(4)
    wgt = torch.rand(16, res[1], res[0], device=DEV)
    out = {}
    for force in (False, True):
        p, n = base_p.clone().requires_grad_(True), base_n.clone().requires_grad_(True)
        group.active_surface_points, group.active_surface_normals = p, n
        tracer = HeliostatRayTracer(scenario, group, blocking_active=True, bitmap_resolution=torch.tensor(res))
        tracer._force_one_cta_per_sample = force
        flux, ic, ot, bl = tracer.trace_rays(inc, mask, tidx)
        (flux * wgt).sum().backward()
        out[force] = (flux.detach(), ic, ot, bl, p.grad, n.grad)
    assert float(out[True][3].min()) < 0.95 and out[True][0].sum() > 0
    for a, b in zip(out[True][:4], out[False][:4]):
        assert torch.equal(a, b)
    for a, b, name in ((out[True][4], out[False][4], "points"), (out[True][5], out[False][5], "normals")):
        assert (a - b).abs().max() <= 2e-5 * b.abs().max(), name
    rows = torch.tensor(HeliostatRayTracer._corner_rows(base_p.shape[1]), device=DEV)
    ca, cb = out[True][4][:, rows], out[False][4][:, rows]
    assert cb.abs().max() > 0 and (ca - cb).abs().max() <= 1e-4 * cb.abs().max()


@pytest.mark.parametrize("one_cta", [False, True])
def test_blocking_with_the_sampler_sharding_adds_up_to_the_unsharded_trace(one_cta):
    """``HeliostatRayTracer(world_size=2, rank=r)`` with blocking on: each rank traces its rows (the blockers are ALL
    heliostats on every rank), rows of the other rank are zero, and the two ranks together give exactly the unsharded
    bitmaps, factors and - summed - gradients (incl. the blockers' corner rows)."""
    from artist_b200 import HeliostatRayTracer

    res = (64, 64)
    ft, scenario, group, mask, tidx, inc = _scene(n=9, ppf=(12, 12), rays=6)
    base_p, base_n = group.active_surface_points.detach().clone(), group.active_surface_normals.detach().clone()
    torch.manual_seed(6)
    wgt = torch.rand(9, res[1], res[0], device=DEV)

    def run(world, rank):
        p, n = base_p.clone().requires_grad_(True), base_n.clone().requires_grad_(True)
        group.active_surface_points, group.active_surface_normals = p, n
        tracer = HeliostatRayTracer(scenario, group, blocking_active=True, world_size=world, rank=rank,
                                    bitmap_resolution=torch.tensor(res))
        tracer._force_one_cta_per_sample = one_cta
        flux, ic, ot, bl = tracer.trace_rays(inc, mask, tidx)
        (flux * wgt).sum().backward()
        return tracer.get_sampler_indices().tolist(), flux.detach(), ic, ot, bl, p.grad, n.grad

    _, full, fic, fot, fbl, gp, gn = run(1, 0)
    parts = [run(2, r) for r in (0, 1)]
    assert sorted(parts[0][0] + parts[1][0]) == list(range(9))
    for rows, flux, ic, ot, bl, _, _ in parts:
        other = [k for k in range(9) if k not in rows]
        assert torch.equal(flux[rows], full[rows]) and flux[other].abs().max() == 0
        assert torch.equal(ic[rows], fic[rows]) and torch.equal(ot[rows], fot[rows]) and torch.equal(bl[rows], fbl[rows])
    for full_g, a, b in ((gp, parts[0][5], parts[1][5]), (gn, parts[0][6], parts[1][6])):
        assert (a + b - full_g).abs().max() <= 2e-5 * full_g.abs().max()
