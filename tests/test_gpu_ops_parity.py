"""-m gpu: NURBS, kinematics, alignment and per-target reduction kernels against the CPU oracle."""
import pytest
import torch

from oracle import artist_oracle as O
from tests import cases

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _nurbs_inputs(n=3, cps=(7, 6), ppf=(11, 9), bump=0.004, seed=2):
    from artist_b200.scenario.synthetic import synthetic_field_tensors

    ft = synthetic_field_tensors(n, control_points=cps, surface_bump=bump, seed=seed)
    ev = O.nurbs_evaluation_grid(*ppf)[None, None].expand(n, 4, -1, -1)
    return ft, ev


@pytest.mark.parametrize("canting", [True, False])
@pytest.mark.parametrize("cps,ppf", [((7, 6), (11, 9)), ((10, 10), (50, 50)), ((20, 20), (13, 13))])
def test_nurbs_forward_bit_equal_points(canting, cps, ppf):
    from artist_b200 import NURBSSurfaces

    ft, ev = _nurbs_inputs(cps=cps, ppf=ppf)
    cant = ft["canting"] if canting else None
    tr = ft["facet_translations"] if canting else None
    pts, nrm = O.nurbs_points_and_normals(ft["nurbs_control_points"], 3, 3, ev, cant, tr)
    surf = NURBSSurfaces(ft["nurbs_degrees"], ft["nurbs_control_points"].to(DEV), device=torch.device(DEV))
    gp, gn = surf.calculate_surface_points_and_normals(ev.to(DEV), None if cant is None else cant.to(DEV),
                                                       None if tr is None else tr.to(DEV))
    # points AND normals follow the reference's op order exactly (the cross product with torch's contraction,
    # the norm as torch's FMA chain: tests/test_oracle_kat.py::test_cpu_cross_product_rounding_the_kernels_reproduce)
    assert torch.equal(gp.cpu(), pts), f"max diff {(gp.cpu() - pts).abs().max():.3e}"
    assert torch.equal(gn.cpu(), nrm), f"max diff {(gn.cpu() - nrm).abs().max():.3e}"


@pytest.mark.parametrize("cps,ppf,deg,canting", [
    ((10, 10), (50, 50), (3, 3), True),
    ((20, 20), (13, 11), (3, 3), True),
    ((5, 6), (40, 64), (2, 3), False),
    ((6, 5), (9, 128), (3, 2), True),
])
def test_nurbs_forward_column_walk_equals_row_table_kernel(cps, ppf, deg, canting, monkeypatch):
    """The column-walk forward (default for shared sorted grids) issues the reference's strict mul/add sequence as
    packed operations: points AND normals must equal the row-table kernel's bit for bit (AB200_NURBS_FWD_ROWTABLE=1
    selects that one at call time), and the points must equal the oracle's."""
    from artist_b200 import ops

    ft, ev = _nurbs_inputs(n=3, cps=cps, ppf=ppf)
    cant = ft["canting"] if canting else None
    tr = ft["facet_translations"] if canting else None
    ku, kv = O.uniform_knots(cps[0], deg[0]).to(DEV), O.uniform_knots(cps[1], deg[1]).to(DEV)
    res = {}
    for mode in ("cols", "rowtable"):
        if mode == "rowtable":
            monkeypatch.setenv("AB200_NURBS_FWD_ROWTABLE", "1")
        else:
            monkeypatch.delenv("AB200_NURBS_FWD_ROWTABLE", raising=False)
        gp, gn = ops.nurbs_points_and_normals(ft["nurbs_control_points"].to(DEV), ev.to(DEV), ku, kv, deg[0], deg[1],
                                              None if cant is None else cant.to(DEV), None if tr is None else tr.to(DEV))
        res[mode] = (gp.cpu(), gn.cpu())
    assert torch.equal(res["cols"][0], res["rowtable"][0]) and torch.equal(res["cols"][1], res["rowtable"][1])
    pts, nrm = O.nurbs_points_and_normals(ft["nurbs_control_points"], deg[0], deg[1], ev, cant, tr)
    assert torch.equal(res["cols"][0], pts), f"max diff {(res['cols'][0] - pts).abs().max():.3e}"
    assert torch.equal(res["cols"][1], nrm), f"max diff {(res['cols'][1] - nrm).abs().max():.3e}"


def test_nurbs_degree_2_and_shared_grid():
    from artist_b200 import ops

    ft, ev = _nurbs_inputs(n=2, cps=(5, 6), ppf=(8, 8))
    pts, nrm = O.nurbs_points_and_normals(ft["nurbs_control_points"], 2, 3, ev, ft["canting"], ft["facet_translations"])
    cp = ft["nurbs_control_points"].to(DEV)
    gp, gn = ops.nurbs_points_and_normals(cp, ev.to(DEV), O.uniform_knots(5, 2).to(DEV), O.uniform_knots(6, 3).to(DEV), 2, 3,
                                          ft["canting"].to(DEV), ft["facet_translations"].to(DEV))
    assert (gp.cpu() - pts).abs().max() <= 1e-6 and (gn.cpu() - nrm).abs().max() <= 1e-6


def test_nurbs_backward_matches_autograd():
    from artist_b200 import NURBSSurfaces

    ft, ev = _nurbs_inputs(n=3, cps=(6, 7), ppf=(12, 10))
    torch.manual_seed(0)
    wp, wn = torch.randn(3, 4, 120, 4), torch.randn(3, 4, 120, 4)
    cp = ft["nurbs_control_points"].clone().requires_grad_(True)
    pts, nrm = O.nurbs_points_and_normals(cp, 3, 3, ev, ft["canting"], ft["facet_translations"])
    ((pts * wp).sum() + (nrm * wn).sum()).backward()
    cpc = ft["nurbs_control_points"].to(DEV).requires_grad_(True)
    surf = NURBSSurfaces(ft["nurbs_degrees"], cpc, device=torch.device(DEV))
    gp, gn = surf.calculate_surface_points_and_normals(ev.to(DEV), ft["canting"].to(DEV), ft["facet_translations"].to(DEV))
    ((gp * wp.to(DEV)).sum() + (gn * wn.to(DEV)).sum()).backward()
    scale = cp.grad.abs().max()
    assert (cpc.grad.cpu() - cp.grad).abs().max() <= 2e-5 * scale


@pytest.mark.parametrize("cps,ppf,deg,canting", [
    ((10, 10), (50, 50), (3, 3), True),     # the benchmark shape
    ((20, 20), (13, 11), (3, 3), True),     # more u-spans than grid rows: the window shifts several rows per step
    ((5, 6), (40, 64), (2, 3), False),      # mixed degrees, no canting, 64 columns per facet
    ((6, 5), (9, 128), (3, 2), True),       # 128 columns: two facets per CTA
])
def test_nurbs_backward_column_walk_kernel(cps, ppf, deg, canting, monkeypatch):
    """The column-walk backward (default for shared sorted grids) against the oracle's autograd and against the
    row-block kernel it replaces (AB200_NURBS_BWD_ROWBLOCK=1 selects that one at call time)."""
    from artist_b200 import ops

    n = 3
    ft, ev = _nurbs_inputs(n=n, cps=cps, ppf=ppf)
    k = ppf[0] * ppf[1]
    torch.manual_seed(1)
    wp, wn = torch.randn(n, 4, k, 4), torch.randn(n, 4, k, 4)
    cant = ft["canting"] if canting else None
    tr = ft["facet_translations"] if canting else None
    cp = ft["nurbs_control_points"].clone().requires_grad_(True)
    pts, nrm = O.nurbs_points_and_normals(cp, deg[0], deg[1], ev, cant, tr)
    ((pts * wp).sum() + (nrm * wn).sum()).backward()
    ku, kv = O.uniform_knots(cps[0], deg[0]).to(DEV), O.uniform_knots(cps[1], deg[1]).to(DEV)
    grads = {}
    for mode in ("cols", "rowblock"):
        if mode == "rowblock":
            monkeypatch.setenv("AB200_NURBS_BWD_ROWBLOCK", "1")
        else:
            monkeypatch.delenv("AB200_NURBS_BWD_ROWBLOCK", raising=False)
        cpc = ft["nurbs_control_points"].to(DEV).requires_grad_(True)
        gp, gn = ops.nurbs_points_and_normals(cpc, ev.to(DEV), ku, kv, deg[0], deg[1], None if cant is None else cant.to(DEV),
                                              None if tr is None else tr.to(DEV))
        ((gp * wp.to(DEV)).sum() + (gn * wn.to(DEV)).sum()).backward()
        grads[mode] = cpc.grad.cpu()
    scale = cp.grad.abs().max()
    assert (grads["cols"] - cp.grad).abs().max() <= 2e-5 * scale
    assert (grads["cols"] - grads["rowblock"]).abs().max() <= 2e-5 * scale


def test_nurbs_backward_unstructured_points_use_generic_gather():
    from artist_b200 import ops

    ft, _ = _nurbs_inputs(n=2, cps=(6, 6), ppf=(5, 5))
    torch.manual_seed(4)
    ev = (0.01 + 0.98 * torch.rand(2, 4, 300, 2)).contiguous()
    assert ops.detect_evaluation_grid(ev.to(DEV)) == (0, 0)
    grid = O.nurbs_evaluation_grid(9, 7)[None, None].expand(2, 4, -1, -1)
    assert ops.detect_evaluation_grid(grid.to(DEV)) == (9, 7)
    wp, wn = torch.randn(2, 4, 300, 4), torch.randn(2, 4, 300, 4)
    cp = ft["nurbs_control_points"].clone().requires_grad_(True)
    pts, nrm = O.nurbs_points_and_normals(cp, 3, 3, ev, ft["canting"], ft["facet_translations"])
    ((pts * wp).sum() + (nrm * wn).sum()).backward()
    cpc = ft["nurbs_control_points"].to(DEV).requires_grad_(True)
    gp, gn = ops.nurbs_points_and_normals(cpc, ev.to(DEV), O.uniform_knots(6, 3).to(DEV), O.uniform_knots(6, 3).to(DEV), 3, 3,
                                          ft["canting"].to(DEV), ft["facet_translations"].to(DEV))
    assert (gp.detach().cpu() - pts.detach()).abs().max() <= 1e-6
    ((gp * wp.to(DEV)).sum() + (gn * wn.to(DEV)).sum()).backward()
    assert (cpc.grad.cpu() - cp.grad).abs().max() <= 2e-5 * cp.grad.abs().max()


def _kin_dev(kin):
    f = lambda x: x.to(DEV).float().contiguous()
    return dict(positions=f(kin.positions), trans=f(kin.translation_deviations), rot=f(kin.rotation_deviations),
                non_opt=f(kin.actuator_non_optimizable), opt=f(kin.actuator_optimizable))


@pytest.mark.parametrize("linear", [True, False])
def test_kinematics_forward_and_backward(linear):
    from artist_b200 import ops

    case = cases.make_case(n=7, points_per_facet=(4, 4), rays=1)
    kin = case["kin"]
    kin.linear = linear
    if not linear:
        kin.actuator_non_optimizable = kin.actuator_non_optimizable.clone()
        kin.actuator_non_optimizable[:, 0] = 1.0
        kin.actuator_non_optimizable[:, 2] = -10.0
        kin.actuator_non_optimizable[:, 3] = 10.0
    g = torch.Generator().manual_seed(1)
    motor = (20000 + 30000 * torch.rand(7, 2, generator=g)) if linear else (torch.rand(7, 2, generator=g) - 0.3)
    wgt = torch.randn(7, 4, 4, generator=g)
    m = motor.clone().requires_grad_(True)
    kin.rotation_deviations = kin.rotation_deviations.clone().requires_grad_(True)
    kin.translation_deviations = kin.translation_deviations.clone().requires_grad_(True)
    kin.actuator_optimizable = kin.actuator_optimizable.clone().requires_grad_(True)
    ref = O.motor_positions_to_orientations(kin, m)
    (ref * wgt).sum().backward()
    d = _kin_dev(kin)
    off = O.initial_orientation_offset().reshape(4, 4).to(DEV)
    mc = motor.to(DEV).requires_grad_(True)
    rot, trans, opt = (d[k].detach().requires_grad_(True) for k in ("rot", "trans", "opt"))
    out = ops.kinematics_orientations(mc, rot, trans, opt if linear else None, d["positions"], d["non_opt"], off, linear)
    # the kernels follow torch-CPU's rounding sequence (unfused left-to-right 4x4 products, strict actuator formulas); what
    # is left are last-bit differences of sin / cos / acos / exp (tools/diag_kinematics_parity.py: 256 heliostats, ideal
    # actuators 56 % of the orientations bit-identical, linear 22 %; rotation part within 2e-7, translation within 2e-6)
    diff = (out.detach().cpu() - ref.detach()).abs()
    assert diff[:, :3, :3].max() <= 6e-7 and diff.max() <= 8e-6, f"{diff[:, :3, :3].max():.2e} {diff.max():.2e}"
    (out * wgt.to(DEV)).sum().backward()
    for got, want, name in ((mc.grad, m.grad, "motor"), (rot.grad, kin.rotation_deviations.grad, "rotation dev"),
                            (trans.grad, kin.translation_deviations.grad, "translation dev")):
        scale = want.abs().max().clamp_min(1e-12)
        assert (got.cpu() - want).abs().max() <= 5e-6 * scale, name      # measured 6e-8 .. 2e-7 (round 1 bar: 2e-3)
    if linear:
        want = kin.actuator_optimizable.grad
        assert (opt.grad.cpu() - want).abs().max() <= 5e-6 * want.abs().max()      # measured 7e-7


@pytest.mark.parametrize("linear", [True, False])
def test_alignment_to_incident_rays(linear):
    from artist_b200 import ops

    case = cases.make_case(n=9, points_per_facet=(4, 4), rays=1, target_pattern=(0, 1))
    kin = case["kin"]
    if not linear:
        kin.linear = False
        kin.actuator_non_optimizable = kin.actuator_non_optimizable.clone()
        kin.actuator_non_optimizable[:, 0] = 1.0
        kin.actuator_non_optimizable[:, 2] = -10.0
        kin.actuator_non_optimizable[:, 3] = 10.0
    ori, motor = O.incident_ray_directions_to_orientations(kin, case["incident"], case["aim"])
    d = _kin_dev(kin)
    off = O.initial_orientation_offset().reshape(4, 4).to(DEV)
    got, gm = ops.kinematics_align_incident(case["incident"].to(DEV), case["aim"].to(DEV), d["rot"], d["trans"],
                                            d["opt"] if linear else None, d["positions"], d["non_opt"], off, linear)
    diff = (got.cpu() - ori).abs()
    assert diff[:, :3, :3].max() <= 1e-6 and diff.max() <= 1.6e-5, f"{diff[:, :3, :3].max():.2e} {diff.max():.2e}"
    assert ((gm.cpu() - motor).abs() / motor.abs().clamp_min(1.0)).max() <= 1e-6       # measured 3e-7 (round 1: 2e-4)


def test_align_apply_forward_backward():
    from artist_b200 import ops

    case = cases.make_case(n=5, points_per_facet=(6, 7), rays=1)
    pts, nrm, ori = case["surface_points"], case["surface_normals"], case["orientations"]
    ap, an = O.align_surfaces(pts, nrm, ori)
    gp, gn = ops.align_surfaces(pts.to(DEV), nrm.to(DEV), ori.to(DEV))
    assert torch.equal(gp.cpu(), ap) and torch.equal(gn.cpu(), an)   # FMA chain == CPU GEMM accumulation
    torch.manual_seed(3)
    w1, w2 = torch.randn_like(ap), torch.randn_like(an)
    p, n_, o = pts.clone().requires_grad_(True), nrm.clone().requires_grad_(True), ori.clone().requires_grad_(True)
    a, b = O.align_surfaces(p, n_, o)
    ((a * w1).sum() + (b * w2).sum()).backward()
    pc, nc, oc = (x.to(DEV).requires_grad_(True) for x in (pts, nrm, ori))
    a2, b2 = ops.align_surfaces(pc, nc, oc)
    ((a2 * w1.to(DEV)).sum() + (b2 * w2.to(DEV)).sum()).backward()
    for got, want in ((pc.grad, p.grad), (nc.grad, n_.grad), (oc.grad, o.grad)):
        assert (got.cpu() - want).abs().max() <= 2e-5 * want.abs().max()
    # replicated samples (mask values > 1) through src_row
    rows = torch.tensor([0, 0, 2, 4, 4, 4], dtype=torch.int32)
    ori6 = ori[rows.long()]
    a3, b3 = ops.align_surfaces(pts.to(DEV), nrm.to(DEV), ori6.to(DEV), src_row=rows.to(DEV))
    ra, rb = O.align_surfaces(pts[rows.long()], nrm[rows.long()], ori6)
    assert torch.equal(a3.cpu(), ra) and torch.equal(b3.cpu(), rb)


def test_bitmaps_per_target():
    from artist_b200 import ops

    torch.manual_seed(0)
    bm = torch.rand(11, 37, 41)
    tidx = torch.tensor([0, 2, 2, 1, 0, 3, 3, 3, 0, 2, 1], dtype=torch.int32)
    ref = O.bitmaps_per_target(bm, tidx, 5)
    got = ops.bitmaps_per_target(bm.to(DEV), tidx.to(DEV), 5)
    assert (got.cpu() - ref).abs().max() <= 1e-6
    assert (got[4] == 0).all()
    bm2 = torch.rand(6, 64, 64)
    got2 = ops.bitmaps_per_target(bm2.to(DEV), tidx[:6].to(DEV), 4)
    assert (got2.cpu() - O.bitmaps_per_target(bm2, tidx[:6], 4)).abs().max() <= 1e-6


def test_device_trig_accuracy():
    """Kernel trig vs torch CPU (SLEEF u10) on sun-shape sized angles: never more than 1 ulp apart."""
    from artist_b200 import ops

    torch.manual_seed(0)
    x = torch.randn(1 << 20) * 2.09e-3
    for mode in (0, 2):
        s, c = ops.debug_trig(x.to(DEV), mode)
        ds = (s.cpu() - torch.sin(x)).abs() / torch.sin(x).abs().clamp_min(1e-30)
        dc = (c.cpu() - torch.cos(x)).abs()
        assert ds.max() <= 1.2e-7 and dc.max() <= 6e-8


@pytest.mark.parametrize("b", [8.0, 8.62967, 7.0, 5.41, 6.39, 7.22, 1.0472, 5.229, 3.0, 0.1, 123.456, 1.9999999])
def test_constant_divisor_quotient_is_ieee_exact(b):
    """The 3-instruction quotient used for the per-target divisors equals IEEE division bit for bit."""
    from artist_b200 import ops

    torch.manual_seed(int(b * 1000))
    n = 1 << 24
    a = torch.cat([torch.rand(n // 2, device=DEV) * 2 * b - 0.5 * b,            # the range the kernels see
                   torch.randn(n // 4, device=DEV) * 1e3, torch.randn(n // 4, device=DEV) * 1e-6])
    qf, qi = ops.debug_const_div(a, b)
    bad = (qf != qi).sum().item()
    assert bad == 0, f"{bad} of {a.numel()} quotients differ for b={b}"


def test_range_guarded_division_is_ieee_exact():
    """div_regular (MUFU.RCP + 5 FFMA, no range check) == __fdiv_rn on the domain the fast ray loops guard:
    |numerator| and |denominator| in [1e-18, 1e18]."""
    from artist_b200 import ops

    torch.manual_seed(0)
    n = 1 << 25
    num = torch.randn(n, device=DEV) * 80.0                                  # (c - o) . n_t, tens of metres
    den = -(torch.rand(n, device=DEV) * 0.999 + 1e-3)                        # d . n_t in (-1, -1e-3)
    num[: n // 8] = torch.exp(torch.empty(n // 8, device=DEV).uniform_(-41.0, 41.0)) * torch.sign(num[: n // 8])
    den[: n // 8] = -torch.exp(torch.empty(n // 8, device=DEV).uniform_(-41.0, 41.0))
    qf, qi = ops.debug_div_regular(num, den)
    bad = (qf != qi).sum().item()
    assert bad == 0, f"{bad} of {n} quotients differ"


@pytest.mark.parametrize("n", [5, 1500])
def test_single_launch_alignment_equals_the_multi_kernel_loop(n, monkeypatch):
    """``ab200_kinematics_align_incident``: the cluster kernel that runs all sweeps in one launch (8 CTAs, fields up to 4096 heliostats)
    against the two-kernels-per-sweep loop (``AB200_ALIGN_MULTI_KERNEL``) - same per-heliostat code, identical bits."""
    from artist_b200 import build_synthetic_scenario

    dev = torch.device("cuda:0")

    def run():
        scenario, group = build_synthetic_scenario(n, number_of_rays=2, points_per_facet=(4, 4), device=dev)
        mask, tidx, _ = scenario.index_mapping(group)
        torch.manual_seed(1)
        inc = torch.nn.functional.normalize(torch.tensor([[0.0, 0.9, -0.43, 0.0]], device=dev)
                                            + 0.2 * torch.randn(n, 4, device=dev) * torch.tensor([1.0, 1.0, 1.0, 0.0], device=dev), dim=1)
        group.activate_heliostats(mask)
        aim = scenario.solar_tower.get_centers_of_target_areas(tidx)
        ori = group.kinematics.incident_ray_directions_to_orientations(incident_ray_directions=inc, aim_points=aim)
        return ori.clone(), group.kinematics.active_motor_positions.clone()

    one = run()
    monkeypatch.setenv("AB200_ALIGN_MULTI_KERNEL", "1")
    multi = run()
    assert torch.isfinite(one[0]).all() and one[1].abs().max() > 0
    assert torch.equal(one[0], multi[0]) and torch.equal(one[1], multi[1])
